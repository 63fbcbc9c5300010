// ORBmatcher.cc -- drop-in replacement of R21/src/ORBmatcher.cc: every candidate loop of ORB_SLAM2::ORBmatcher runs on
// the GPU through the C ABI of include/orbcuda.h.
//
// What stays on the host, written against the reference's own Frame / KeyFrame / MapPoint classes:
//   * the geometry in front of each window search (camera-frame transform, projection, image / distance / viewing-angle
//     gates, predicted level) -- a few 3x3 products per point, evaluated with the same cv::Mat expressions as the
//     reference so that no rounding of cv::gemm has to be re-modelled; the result is one orbm_proj_point_t per point;
//   * flattening Frame::mGrid into CSR (a KeyFrame hides its copy of the grid: it is rebuilt on the device from
//     mvKeysUn with the session's image bounds);
//   * the map bookkeeping behind Fuse (Replace / AddObservation), applied in the reference's point order.
// Line references are to R21/src/ORBmatcher.cc.
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <cstring>

#include "ORBmatcher_accel.h"

using namespace std;

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
int ORBmatcher::msDevice = 0;

namespace
{

static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint_t), "cv::KeyPoint must be the 28-byte POD of orbcuda.h");

inline const orb_keypoint_t* kp_ptr(const vector<cv::KeyPoint>& v)
{
    return v.empty() ? 0 : reinterpret_cast<const orb_keypoint_t*>(&v[0]);
}

// The 64 x 48 cell lists of a frame as CSR (cell = ix * 48 + iy, indices in push_back order) + what a window lookup needs.
struct GridView
{
    vector<int32_t> ptr, idx;
    float bounds[4];          // mnMinX, mnMaxX, mnMinY, mnMaxY of the session (Frame statics): cell size
    float origin[2];          // lookup origin of a KeyFrame (its int-truncated mnMinX, mnMinY)
    bool has_origin;
    GridView() : ptr(ORBF_GRID_COLS * ORBF_GRID_ROWS + 1, 0), has_origin(false) {}
    const float* origin_or_null() const { return has_origin ? origin : 0; }
};

void session_bounds(float b[4])
{
    b[0] = Frame::mnMinX; b[1] = Frame::mnMaxX; b[2] = Frame::mnMinY; b[3] = Frame::mnMaxY;
}

// Frame::mGrid is public: copy the lists as they are
GridView grid_of(const Frame& F)
{
    GridView g;
    session_bounds(g.bounds);
    g.idx.reserve(F.mvKeysUn.size() + 1);
    int c = 0;
    for(int ix = 0; ix < ORBF_GRID_COLS; ix++)
        for(int iy = 0; iy < ORBF_GRID_ROWS; iy++, c++)
        {
            g.ptr[c] = (int32_t)g.idx.size();
            const vector<size_t>& cell = F.mGrid[ix][iy];
            for(size_t k = 0; k < cell.size(); k++) g.idx.push_back((int32_t)cell[k]);
        }
    g.ptr[c] = (int32_t)g.idx.size();
    if(g.idx.empty()) g.idx.push_back(0);
    return g;
}

// KeyFrame::mGrid is protected (a copy of the Frame's lists, KeyFrame.cc:44-50): rebuilt by Frame::AssignFeaturesToGrid's
// device twin from the same key points and the same (session) bounds.  Lookups start from the key frame's int bounds.
GridView grid_of(KeyFrame* pKF)
{
    GridView g;
    session_bounds(g.bounds);
    g.idx.assign(pKF->mvKeysUn.size() + 1, 0);
    int n = 0;
    orbaccel::check(orbf_assign_grid(kp_ptr(pKF->mvKeysUn), (int)pKF->mvKeysUn.size(), g.bounds, &g.ptr[0], &g.idx[0], &n,
                                     ORBmatcher::msDevice), "orbf_assign_grid");
    g.origin[0] = (float)pKF->mnMinX; g.origin[1] = (float)pKF->mnMinY;
    g.has_origin = true;
    return g;
}

inline void copy_desc(unsigned char* dst, const cv::Mat& row) { memcpy(dst, row.ptr(0), 32); }

}  // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
{
    return orb_hamming256(a.ptr(0), b.ptr(0));
}

// ---- :45-130 ------------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
{
    const int nMP = (int)vpMapPoints.size(), N = (int)F.mvKeysUn.size();
    if(nMP == 0 || N == 0) return 0;
    const GridView g = grid_of(F);
    vector<orbm_map_point_view_t> mp(nMP);
    vector<unsigned char> desc((size_t)nMP * 32, 0), occ(N);
    for(int i = 0; i < nMP; i++)
    {
        MapPoint* p = vpMapPoints[i];
        mp[i].in_view = p->mbTrackInView && !p->isBad();                      // :52-56
        mp[i].proj_x = p->mTrackProjX; mp[i].proj_y = p->mTrackProjY; mp[i].proj_xr = p->mTrackProjXR;
        mp[i].view_cos = p->mTrackViewCos; mp[i].level = p->mnTrackScaleLevel;
        mp[i].obs_positive = p->Observations() > 0;
        if(mp[i].in_view) copy_desc(&desc[(size_t)i * 32], p->GetDescriptor());
    }
    for(int f = 0; f < N; f++) occ[f] = F.mvpMapPoints[f] && F.mvpMapPoints[f]->Observations() > 0;   // :82-84
    vector<int32_t> fp(N), pf(nMP);
    int nmatches = 0;
    orbaccel::check(orbm_search_by_projection_frame(kp_ptr(F.mvKeysUn), F.mDescriptors.ptr(0), &F.mvuRight[0], &occ[0], N, &g.ptr[0], &g.idx[0],
                                                    g.bounds, &F.mvScaleFactors[0], (int)F.mvScaleFactors.size(), &mp[0], &desc[0], nMP, th,
                                                    mfNNratio, TH_HIGH, &fp[0], &pf[0], &nmatches, msDevice),
                    "orbm_search_by_projection_frame");
    for(int f = 0; f < N; f++)
        if(fp[f] >= 0) F.mvpMapPoints[f] = vpMapPoints[fp[f]];                // :121
    return nmatches;
}

// ---- :290-403 -----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
{
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
    // similarity -> rotation, translation, camera centre (:299-303)
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;

    set<MapPoint*> found(vpMatched.begin(), vpMatched.end());
    found.erase(static_cast<MapPoint*>(NULL));

    const int nP = (int)vpPoints.size(), N = pKF->N;
    if(nP == 0 || N == 0) return 0;
    vector<orbm_proj_point_t> pts(nP);
    vector<unsigned char> desc((size_t)nP * 32, 0), occ(N);
    for(int i = 0; i < nP; i++)
    {
        orbm_proj_point_t& q = pts[i];
        memset(&q, 0, sizeof(q));
        MapPoint* pMP = vpPoints[i];
        if(pMP->isBad() || found.count(pMP)) continue;                        // :317-318
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if(p3Dc.at<float>(2) < 0.0) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz;
        const float y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if(!pKF->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist = cv::norm(PO);
        if(dist < minDistance || dist > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if(PO.dot(Pn) < 0.5 * dist) continue;                                 // viewing angle below 60 degrees
        q.u = u; q.v = v; q.octave = pMP->PredictScale(dist, pKF); q.valid = 1;
        copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
    }
    for(int f = 0; f < N; f++) occ[f] = vpMatched[f] != NULL;                 // :370-371
    const GridView g = grid_of(pKF);
    vector<int32_t> fp(N), pf(nP);
    int nmatches = 0;
    orbaccel::check(orbm_search_by_projection_sim3(kp_ptr(pKF->mvKeysUn), pKF->mDescriptors.ptr(0), &occ[0], N, &g.ptr[0], &g.idx[0], g.bounds,
                                                   &pKF->mvScaleFactors[0], (int)pKF->mvScaleFactors.size(), &pts[0], &desc[0], nP, (float)th,
                                                   TH_LOW, &fp[0], &pf[0], &nmatches, g.origin_or_null(), msDevice),
                    "orbm_search_by_projection_sim3");
    for(int f = 0; f < N; f++)
        if(fp[f] >= 0) vpMatched[f] = vpPoints[fp[f]];                        // :381
    return nmatches;
}

// ---- :159-288, :522-655 ------------------------------------------------------------------------------------------
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
{
    return orbaccel::SearchByBoW_KF_F(pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation, msDevice);
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
{
    return orbaccel::SearchByBoW_KF_KF(pKF1, pKF2, vpMatches12, mfNNratio, mbCheckOrientation, msDevice);
}

// ---- :405-520 -----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize)
{
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    vnMatches12 = vector<int>(n1, -1);
    if(n1 == 0) return 0;
    const GridView g = grid_of(F2);
    vector<float> xy((size_t)2 * n1);
    for(int i = 0; i < n1; i++) { xy[2 * i] = vbPrevMatched[i].x; xy[2 * i + 1] = vbPrevMatched[i].y; }
    vector<int32_t> m12(n1);
    int nmatches = 0;
    orbaccel::check(orbm_search_for_initialization(kp_ptr(F1.mvKeysUn), F1.mDescriptors.ptr(0), n1, kp_ptr(F2.mvKeysUn), F2.mDescriptors.ptr(0),
                                                   n2, &g.ptr[0], &g.idx[0], g.bounds, &xy[0], windowSize, mfNNratio, mbCheckOrientation, TH_LOW,
                                                   &m12[0], &nmatches, msDevice), "orbm_search_for_initialization");
    for(int i = 0; i < n1; i++)
    {
        vnMatches12[i] = m12[i];
        vbPrevMatched[i].x = xy[2 * i]; vbPrevMatched[i].y = xy[2 * i + 1];  // :513-516, already applied by the library
    }
    return nmatches;
}

// ---- :657-823 -----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo)
{
    // epipole of camera 1 in image 2 (:664-670)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    float f12[9];
    for(int r = 0; r < 3; r++)
        for(int c = 0; c < 3; c++) f12[3 * r + c] = F12.at<float>(r, c);
    return orbaccel::SearchForTriangulation(pKF1, pKF2, f12, ex, ey, vMatchedPairs, bOnlyStereo, mbCheckOrientation, msDevice);
}

// ---- :825-975 -----------------------------------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)
{
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy, bf = pKF->mbf;
    cv::Mat Ow = pKF->GetCameraCenter();
    const int nMPs = (int)vpMapPoints.size(), N = pKF->N;
    if(nMPs == 0 || N == 0) return 0;
    vector<orbm_proj_point_t> pts(nMPs);
    vector<unsigned char> desc((size_t)nMPs * 32, 0);
    for(int i = 0; i < nMPs; i++)
    {
        orbm_proj_point_t& q = pts[i];
        memset(&q, 0, sizeof(q));
        MapPoint* pMP = vpMapPoints[i];
        if(!pMP) continue;
        if(pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                 // re-checked when the result is applied
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if(p3Dc.at<float>(2) < 0.0f) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz;
        const float y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if(!pKF->IsInImage(u, v)) continue;
        const float ur = u - bf * invz;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if(dist3D < minDistance || dist3D > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if(PO.dot(Pn) < 0.5 * dist3D) continue;
        q.u = u; q.v = v; q.ur = ur; q.octave = pMP->PredictScale(dist3D, pKF); q.valid = 1;
        copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
    }
    const GridView g = grid_of(pKF);
    vector<int32_t> best(nMPs), bestd(nMPs);
    orbaccel::check(orbm_window_best_match(kp_ptr(pKF->mvKeysUn), pKF->mDescriptors.ptr(0), &pKF->mvuRight[0], N, &g.ptr[0], &g.idx[0], g.bounds,
                                           &pKF->mvScaleFactors[0], &pKF->mvInvLevelSigma2[0], (int)pKF->mvScaleFactors.size(), &pts[0],
                                           &desc[0], nMPs, th, &best[0], &bestd[0], g.origin_or_null(), msDevice), "orbm_window_best_match");
    // map updates in the reference's order (:945-970).  An earlier update can only invalidate a later point (Replace marks it
    // bad, AddObservation puts it into pKF); its window search does not depend on earlier points.
    int nFused = 0;
    for(int i = 0; i < nMPs; i++)
    {
        if(!pts[i].valid || bestd[i] > TH_LOW) continue;
        MapPoint* pMP = vpMapPoints[i];
        if(pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if(pMPinKF)
        {
            if(!pMPinKF->isBad())
            {
                if(pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        }
        else
        {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---- :977-1100 ----------------------------------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
{
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const set<MapPoint*> found = pKF->GetMapPoints();
    const int nP = (int)vpPoints.size(), N = pKF->N;
    if(nP == 0 || N == 0) return 0;
    vector<orbm_proj_point_t> pts(nP);
    vector<unsigned char> desc((size_t)nP * 32, 0);
    for(int i = 0; i < nP; i++)
    {
        orbm_proj_point_t& q = pts[i];
        memset(&q, 0, sizeof(q));
        MapPoint* pMP = vpPoints[i];
        if(pMP->isBad() || found.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if(p3Dc.at<float>(2) < 0.0f) continue;
        const float invz = 1.0 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz;
        const float y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if(!pKF->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if(dist3D < minDistance || dist3D > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if(PO.dot(Pn) < 0.5 * dist3D) continue;
        q.u = u; q.v = v; q.octave = pMP->PredictScale(dist3D, pKF); q.valid = 1;
        copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
    }
    const GridView g = grid_of(pKF);
    vector<int32_t> best(nP), bestd(nP);
    orbaccel::check(orbm_window_best_match(kp_ptr(pKF->mvKeysUn), pKF->mDescriptors.ptr(0), 0, N, &g.ptr[0], &g.idx[0], g.bounds,
                                           &pKF->mvScaleFactors[0], 0, (int)pKF->mvScaleFactors.size(), &pts[0], &desc[0], nP, th, &best[0],
                                           &bestd[0], g.origin_or_null(), msDevice), "orbm_window_best_match");
    int nFused = 0;
    for(int i = 0; i < nP; i++)                                               // :1075-1095
    {
        if(!pts[i].valid || bestd[i] > TH_LOW) continue;
        MapPoint* pMP = vpPoints[i];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if(pMPinKF)
        {
            if(!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        }
        else
        {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---- :1102-1326 ---------------------------------------------------------------------------------------------------
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th)
{
    const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;                                                 // :1119-1121
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;

    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = (int)vpMapPoints2.size();
    vector<bool> done1(N1, false), done2(N2, false);
    for(int i = 0; i < N1; i++)                                               // :1132-1142
    {
        MapPoint* pMP = vpMatches12[i];
        if(pMP)
        {
            done1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if(idx2 >= 0 && idx2 < N2) done2[idx2] = true;
        }
    }

    // one pass per direction: points of `from` (already in its camera frame through Rw, tw) into image `to`
    struct Pass
    {
        static void run(const vector<MapPoint*>& vp, const vector<bool>& done, const cv::Mat& Rw, const cv::Mat& tw, const cv::Mat& sR,
                        const cv::Mat& t, KeyFrame* to, float fx, float fy, float cx, float cy, float th, vector<int>& match)
        {
            const int n = (int)vp.size(), N = to->N;
            match.assign(n, -1);
            if(n == 0 || N == 0) return;
            vector<orbm_proj_point_t> pts(n);
            vector<unsigned char> desc((size_t)n * 32, 0);
            for(int i = 0; i < n; i++)
            {
                orbm_proj_point_t& q = pts[i];
                memset(&q, 0, sizeof(q));
                MapPoint* pMP = vp[i];
                if(!pMP || done[i]) continue;
                if(pMP->isBad()) continue;
                cv::Mat p3Dw = pMP->GetWorldPos();
                cv::Mat p3Da = Rw * p3Dw + tw;
                cv::Mat p3Db = sR * p3Da + t;
                if(p3Db.at<float>(2) < 0.0) continue;
                const float invz = 1.0 / p3Db.at<float>(2);
                const float x = p3Db.at<float>(0) * invz;
                const float y = p3Db.at<float>(1) * invz;
                const float u = fx * x + cx;
                const float v = fy * y + cy;
                if(!to->IsInImage(u, v)) continue;
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                const float dist3D = cv::norm(p3Db);
                if(dist3D < minDistance || dist3D > maxDistance) continue;
                q.u = u; q.v = v; q.octave = pMP->PredictScale(dist3D, to); q.valid = 1;
                copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
            }
            const GridView g = grid_of(to);
            vector<int32_t> best(n), bestd(n);
            orbaccel::check(orbm_window_best_match(kp_ptr(to->mvKeysUn), to->mDescriptors.ptr(0), 0, N, &g.ptr[0], &g.idx[0], g.bounds,
                                                   &to->mvScaleFactors[0], 0, (int)to->mvScaleFactors.size(), &pts[0], &desc[0], n, th, &best[0],
                                                   &bestd[0], g.origin_or_null(), ORBmatcher::msDevice), "orbm_window_best_match");
            for(int i = 0; i < n; i++)
                if(pts[i].valid && bestd[i] <= ORBmatcher::TH_HIGH) match[i] = best[i];
        }
    };
    vector<int> vnMatch1, vnMatch2;
    Pass::run(vpMapPoints1, done1, R1w, t1w, sR21, t21, pKF2, fx, fy, cx, cy, th, vnMatch1);     // KF1 -> KF2 (:1148-1225)
    Pass::run(vpMapPoints2, done2, R2w, t2w, sR12, t12, pKF1, fx, fy, cx, cy, th, vnMatch2);     // KF2 -> KF1 (:1228-1305)

    int nFound = 0;
    for(int i1 = 0; i1 < N1; i1++)                                            // mutual agreement (:1308-1323)
    {
        const int idx2 = vnMatch1[i1];
        if(idx2 >= 0 && vnMatch2[idx2] == i1)
        {
            vpMatches12[i1] = vpMapPoints2[idx2];
            nFound++;
        }
    }
    return nFound;
}

// ---- :1328-1470 ---------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;      // :1348-1349
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

    const int nP = LastFrame.N, N = (int)CurrentFrame.mvKeysUn.size();
    if(nP == 0 || N == 0) return 0;
    vector<orbm_proj_point_t> pts(nP);
    vector<unsigned char> desc((size_t)nP * 32, 0), occ(N);
    for(int i = 0; i < nP; i++)
    {
        orbm_proj_point_t& q = pts[i];
        memset(&q, 0, sizeof(q));
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if(!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        if(invzc < 0) continue;
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if(u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if(v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        q.u = u; q.v = v; q.ur = u - CurrentFrame.mbf * invzc;               // :1410
        q.octave = LastFrame.mvKeys[i].octave;                                // :1378
        q.angle = LastFrame.mvKeysUn[i].angle;
        q.obs_positive = pMP->Observations() > 0;
        q.valid = 1;
        copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
    }
    for(int f = 0; f < N; f++) occ[f] = CurrentFrame.mvpMapPoints[f] && CurrentFrame.mvpMapPoints[f]->Observations() > 0;   // :1404-1406
    const GridView g = grid_of(CurrentFrame);
    vector<int32_t> fp(N), pf(nP);
    int nmatches = 0;
    orbaccel::check(orbm_search_by_projection_last_frame(kp_ptr(CurrentFrame.mvKeysUn), CurrentFrame.mDescriptors.ptr(0), &CurrentFrame.mvuRight[0],
                                                         &occ[0], N, &g.ptr[0], &g.idx[0], g.bounds, &CurrentFrame.mvScaleFactors[0],
                                                         (int)CurrentFrame.mvScaleFactors.size(), &pts[0], &desc[0], nP, th,
                                                         bForward ? 1 : (bBackward ? 2 : 0), mbCheckOrientation, TH_HIGH, &fp[0], &pf[0],
                                                         &nmatches, msDevice), "orbm_search_by_projection_last_frame");
    for(int f = 0; f < N; f++)
    {
        if(fp[f] >= 0) CurrentFrame.mvpMapPoints[f] = LastFrame.mvpMapPoints[fp[f]];          // :1430
        else if(fp[f] == -2) CurrentFrame.mvpMapPoints[f] = static_cast<MapPoint*>(NULL);      // :1459
    }
    return nmatches;
}

// ---- :1472-1599 ---------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int nP = (int)vpMPs.size(), N = (int)CurrentFrame.mvKeysUn.size();
    if(nP == 0 || N == 0) return 0;
    vector<orbm_proj_point_t> pts(nP);
    vector<unsigned char> desc((size_t)nP * 32, 0), occ(N);
    for(int i = 0; i < nP; i++)
    {
        orbm_proj_point_t& q = pts[i];
        memset(&q, 0, sizeof(q));
        MapPoint* pMP = vpMPs[i];
        if(!pMP) continue;
        if(pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if(u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
        if(v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        cv::Mat PO = x3Dw - Ow;
        const float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if(dist3D < minDistance || dist3D > maxDistance) continue;
        q.u = u; q.v = v; q.octave = pMP->PredictScale(dist3D, &CurrentFrame);               // :1522
        q.angle = pKF->mvKeysUn[i].angle;
        q.valid = 1;
        copy_desc(&desc[(size_t)i * 32], pMP->GetDescriptor());
    }
    for(int f = 0; f < N; f++) occ[f] = CurrentFrame.mvpMapPoints[f] != NULL;                 // :1543-1544
    const GridView g = grid_of(CurrentFrame);
    vector<int32_t> fp(N), pf(nP);
    int nmatches = 0;
    orbaccel::check(orbm_search_by_projection_keyframe(kp_ptr(CurrentFrame.mvKeysUn), CurrentFrame.mDescriptors.ptr(0), &occ[0], N, &g.ptr[0],
                                                       &g.idx[0], g.bounds, &CurrentFrame.mvScaleFactors[0],
                                                       (int)CurrentFrame.mvScaleFactors.size(), &pts[0], &desc[0], nP, th, ORBdist,
                                                       mbCheckOrientation, &fp[0], &pf[0], &nmatches, msDevice),
                    "orbm_search_by_projection_keyframe");
    for(int f = 0; f < N; f++)
    {
        if(fp[f] >= 0) CurrentFrame.mvpMapPoints[f] = vpMPs[fp[f]];                           // :1559
        else if(fp[f] == -2) CurrentFrame.mvpMapPoints[f] = NULL;                             // :1588
    }
    return nmatches;
}

}  // namespace ORB_SLAM2
