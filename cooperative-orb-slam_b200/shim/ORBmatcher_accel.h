// ORBmatcher_accel.h -- adapters that route the reference's ORBmatcher hot loops to liborbcuda.
//
// The reference's ORBmatcher methods take Frame / KeyFrame / MapPoint objects (R21/include/ORBmatcher.h).
// Those classes stay untouched; the function templates below are duck-typed on exactly the members the
// reference's loops read, flatten them into the POD arrays of include/orbcuda.h, call the C ABI and
// scatter the result back in the reference's own output type.  Inside ORBmatcher.cc a method body becomes
// one call, e.g.
//     int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
//     { return orbaccel::SearchByBoW_KF_F(pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation); }
// ORBmatcher::DescriptorDistance (R21/src/ORBmatcher.cc:1647) -> orb_hamming256(a.ptr(), b.ptr()).
#ifndef ORBMATCHER_ACCEL_H
#define ORBMATCHER_ACCEL_H

#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "orbcuda.h"

namespace orbaccel
{

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>) -> CSR
template <class FeatureVector>
struct FlatFeatVec
{
    std::vector<int32_t> ids, ptr, idx;
    orbm_featvec_t view;
    explicit FlatFeatVec(const FeatureVector& fv)
    {
        ptr.push_back(0);
        for(typename FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        {
            ids.push_back((int32_t)it->first);
            for(size_t k = 0; k < it->second.size(); k++)
                idx.push_back((int32_t)it->second[k]);
            ptr.push_back((int32_t)idx.size());
        }
        view.n_nodes = (int32_t)ids.size();
        view.node_ids = ids.empty() ? 0 : &ids[0];
        view.ptr = &ptr[0];
        view.idx = idx.empty() ? 0 : &idx[0];
    }
};

inline void check(int rc, const char* what)
{
    if(rc != ORB_OK)
        throw std::runtime_error(std::string(what) + " failed: " + orb_last_error());
}

template <class KeyPointVec>
inline std::vector<float> angles(const KeyPointVec& v)
{
    std::vector<float> a(v.size());
    for(size_t i = 0; i < v.size(); i++) a[i] = v[i].angle;
    return a;
}

// R21/src/ORBmatcher.cc:159-288.  mDescriptors must be continuous N x 32 CV_8U (it is: ORBextractor output).
template <class KeyFrame, class Frame, class MapPoint>
int SearchByBoW_KF_F(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches, float nnratio, bool checkOri,
                     int device = 0)
{
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    const int nKF = (int)vpMapPointsKF.size();
    std::vector<unsigned char> valid(nKF);
    for(int i = 0; i < nKF; i++) valid[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    FlatFeatVec<typename std::remove_reference<decltype(pKF->mFeatVec)>::type> fk(pKF->mFeatVec);
    FlatFeatVec<typename std::remove_reference<decltype(F.mFeatVec)>::type> ff(F.mFeatVec);
    std::vector<float> aKF = angles(pKF->mvKeysUn), aF = angles(F.mvKeys);
    std::vector<int32_t> match(F.N);
    int nmatches = 0;
    check(orbm_search_by_bow_kf_f(pKF->mDescriptors.ptr(0), &aKF[0], &valid[0], nKF, &fk.view, F.mDescriptors.ptr(0), &aF[0], F.N,
                                  &ff.view, nnratio, checkOri, &match[0], &nmatches, device), "orbm_search_by_bow_kf_f");
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    for(int j = 0; j < F.N; j++)
        if(match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];
    return nmatches;
}

// R21/src/ORBmatcher.cc:522-655
template <class KeyFrame, class MapPoint>
int SearchByBoW_KF_KF(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, float nnratio, bool checkOri,
                      int device = 0)
{
    const std::vector<MapPoint*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    const int n1 = (int)mp1.size(), n2 = (int)mp2.size();
    std::vector<unsigned char> v1(n1), v2(n2);
    for(int i = 0; i < n1; i++) v1[i] = mp1[i] && !mp1[i]->isBad();
    for(int i = 0; i < n2; i++) v2[i] = mp2[i] && !mp2[i]->isBad();
    FlatFeatVec<typename std::remove_reference<decltype(pKF1->mFeatVec)>::type> f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    std::vector<float> a1 = angles(pKF1->mvKeysUn), a2 = angles(pKF2->mvKeysUn);
    std::vector<int32_t> match(n1);
    int nmatches = 0;
    check(orbm_search_by_bow_kf_kf(pKF1->mDescriptors.ptr(0), &a1[0], &v1[0], n1, &f1.view, pKF2->mDescriptors.ptr(0), &a2[0], &v2[0],
                                   n2, &f2.view, nnratio, checkOri, &match[0], &nmatches, device), "orbm_search_by_bow_kf_kf");
    vpMatches12 = std::vector<MapPoint*>(n1, static_cast<MapPoint*>(NULL));
    for(int i = 0; i < n1; i++)
        if(match[i] >= 0) vpMatches12[i] = mp2[match[i]];
    return nmatches;
}

// R21/src/ORBmatcher.cc:657-823.  F12 is a 3x3 CV_32F cv::Mat; the epipole (:664-670) is computed by the caller's
// own geometry code and passed in.
template <class KeyFrame>
int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, const float* F12_rowmajor, float ex, float ey,
                           std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, bool checkOri,
                           int device = 0)
{
    const int n1 = pKF1->N, n2 = pKF2->N;
    std::vector<orbm_tri_feature_t> t1(n1), t2(n2);
    for(int i = 0; i < n1; i++)
    {
        t1[i].x = pKF1->mvKeysUn[i].pt.x; t1[i].y = pKF1->mvKeysUn[i].pt.y; t1[i].angle = pKF1->mvKeysUn[i].angle;
        t1[i].octave = pKF1->mvKeysUn[i].octave; t1[i].u_right = pKF1->mvuRight[i]; t1[i].has_mp = pKF1->GetMapPoint(i) ? 1 : 0;
    }
    for(int i = 0; i < n2; i++)
    {
        t2[i].x = pKF2->mvKeysUn[i].pt.x; t2[i].y = pKF2->mvKeysUn[i].pt.y; t2[i].angle = pKF2->mvKeysUn[i].angle;
        t2[i].octave = pKF2->mvKeysUn[i].octave; t2[i].u_right = pKF2->mvuRight[i]; t2[i].has_mp = pKF2->GetMapPoint(i) ? 1 : 0;
    }
    FlatFeatVec<typename std::remove_reference<decltype(pKF1->mFeatVec)>::type> f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    std::vector<int32_t> pairs((size_t)2 * (n1 > 0 ? n1 : 1));
    int nmatches = 0;
    check(orbm_search_for_triangulation(pKF1->mDescriptors.ptr(0), &t1[0], n1, &f1.view, pKF2->mDescriptors.ptr(0), &t2[0], n2,
                                        &f2.view, F12_rowmajor, ex, ey, &pKF2->mvScaleFactors[0], &pKF2->mvLevelSigma2[0],
                                        bOnlyStereo, checkOri, &pairs[0], n1, &nmatches, device), "orbm_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for(int k = 0; k < nmatches; k++)
        vMatchedPairs.push_back(std::make_pair((size_t)pairs[2 * k], (size_t)pairs[2 * k + 1]));
    return nmatches;
}

// Frame::ComputeStereoMatches (R21/src/Frame.cc:471-645): body becomes
//     orbaccel::ComputeStereoMatches(*this);
template <class Frame>
void ComputeStereoMatches(Frame& F)
{
    F.mvuRight = std::vector<float>(F.N, -1.0f);
    F.mvDepth = std::vector<float>(F.N, -1.0f);
    const int Nr = (int)F.mvKeysRight.size();
    if(F.N == 0 || Nr == 0) return;
    static_assert(sizeof(F.mvKeys[0]) == sizeof(orb_keypoint_t), "cv::KeyPoint layout");
    int n = 0;
    // mb is mbf/fx; the reference reads it before assigning it (R21/src/Frame.cc:116 vs :501)
    check(orbm_stereo_matches(F.mpORBextractorLeft->Handle(), F.mpORBextractorRight->Handle(),
                              reinterpret_cast<const orb_keypoint_t*>(&F.mvKeys[0]), F.mDescriptors.ptr(0), F.N,
                              reinterpret_cast<const orb_keypoint_t*>(&F.mvKeysRight[0]), F.mDescriptorsRight.ptr(0), Nr,
                              F.mbf, F.mbf / F.fx, &F.mvuRight[0], &F.mvDepth[0], &n), "orbm_stereo_matches");
}

// ---------------------------------------------------------------------------------------------------------------
// Frame side (R21/src/Frame.cc) and the projection searches that consume it.
// The reference keeps mGrid[64][48] of std::vector<size_t>; the device path keeps the same lists as CSR.
struct FrameGrid
{
    std::vector<int32_t> cell_ptr, cell_idx;
    float bounds[4];                       // mnMinX, mnMaxX, mnMinY, mnMaxY
    FrameGrid() : cell_ptr(ORBF_GRID_COLS * ORBF_GRID_ROWS + 1, 0) { bounds[0] = bounds[1] = bounds[2] = bounds[3] = 0.f; }
};

// void Frame::UndistortKeyPoints()  :409-439  -- body becomes  orbaccel::UndistortKeyPoints(*this);
template <class Frame>
void UndistortKeyPoints(Frame& F, int device = 0)
{
    static_assert(sizeof(F.mvKeys[0]) == sizeof(orb_keypoint_t), "cv::KeyPoint layout");
    F.mvKeysUn.resize(F.mvKeys.size());
    if(F.mvKeys.empty()) return;
    const float K[4] = {F.fx, F.fy, F.cx, F.cy};
    check(orbf_undistort_keypoints(reinterpret_cast<const orb_keypoint_t*>(&F.mvKeys[0]), (int)F.mvKeys.size(), K,
                                   F.mDistCoef.template ptr<float>(0), F.mDistCoef.rows,
                                   reinterpret_cast<orb_keypoint_t*>(&F.mvKeysUn[0]), device), "orbf_undistort_keypoints");
}

// void Frame::ComputeImageBounds(imLeft) :441-470 + AssignFeaturesToGrid() :235-250
template <class Frame>
FrameGrid AssignFeaturesToGrid(const Frame& F, int cols, int rows, int device = 0)
{
    FrameGrid g;
    const float K[4] = {F.fx, F.fy, F.cx, F.cy};
    check(orbf_image_bounds(cols, rows, K, F.mDistCoef.template ptr<float>(0), F.mDistCoef.rows, g.bounds, device), "orbf_image_bounds");
    g.cell_idx.resize(F.mvKeysUn.size() ? F.mvKeysUn.size() : 1);
    int n = 0;
    check(orbf_assign_grid(F.mvKeysUn.empty() ? 0 : reinterpret_cast<const orb_keypoint_t*>(&F.mvKeysUn[0]), (int)F.mvKeysUn.size(),
                           g.bounds, &g.cell_ptr[0], &g.cell_idx[0], &n, device), "orbf_assign_grid");
    return g;
}

// UndistortKeyPoints() :409-439 and AssignFeaturesToGrid() :235-250 of the Frame constructors in ONE pass over the device
// (orbf_build_frame: one upload of mvKeys, two kernels, one download).  `bounds` is the camera's image bounds -- the reference
// computes them once (mbInitialComputations, :96-110): keep the FrameGrid::bounds of the first frame (AssignFeaturesToGrid above)
// and pass them to every later frame.
template <class Frame>
FrameGrid BuildFrame(Frame& F, const float bounds[4], int device = 0)
{
    static_assert(sizeof(F.mvKeys[0]) == sizeof(orb_keypoint_t), "cv::KeyPoint layout");
    FrameGrid g;
    for(int i = 0; i < 4; i++) g.bounds[i] = bounds[i];
    F.mvKeysUn.resize(F.mvKeys.size());
    g.cell_idx.resize(F.mvKeys.size() ? F.mvKeys.size() : 1);
    const float K[4] = {F.fx, F.fy, F.cx, F.cy};
    int n = 0;
    check(orbf_build_frame(F.mvKeys.empty() ? 0 : reinterpret_cast<const orb_keypoint_t*>(&F.mvKeys[0]), (int)F.mvKeys.size(), K,
                           F.mDistCoef.template ptr<float>(0), F.mDistCoef.rows, g.bounds,
                           F.mvKeysUn.empty() ? 0 : reinterpret_cast<orb_keypoint_t*>(&F.mvKeysUn[0]), &g.cell_ptr[0], &g.cell_idx[0], &n, device),
          "orbf_build_frame");
    return g;
}

// int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)  :45-130
template <class Frame, class MapPoint>
int SearchByProjection(Frame& F, const FrameGrid& grid, const std::vector<MapPoint*>& vpMapPoints, float th, float nnratio,
                       int th_high = 100, int device = 0)
{
    const int nMP = (int)vpMapPoints.size(), N = (int)F.mvKeysUn.size();
    if(nMP == 0 || N == 0) return 0;
    std::vector<orbm_map_point_view_t> mp(nMP);
    std::vector<unsigned char> desc((size_t)nMP * 32, 0), occ(N);
    for(int i = 0; i < nMP; i++)
    {
        MapPoint* p = vpMapPoints[i];
        mp[i].in_view = p->mbTrackInView && !p->isBad();                       // :52-56
        mp[i].proj_x = p->mTrackProjX; mp[i].proj_y = p->mTrackProjY; mp[i].proj_xr = p->mTrackProjXR;
        mp[i].view_cos = p->mTrackViewCos; mp[i].level = p->mnTrackScaleLevel;
        mp[i].obs_positive = p->Observations() > 0;
        if(mp[i].in_view) std::memcpy(&desc[(size_t)i * 32], p->GetDescriptor().ptr(0), 32);
    }
    for(int f = 0; f < N; f++) occ[f] = F.mvpMapPoints[f] && F.mvpMapPoints[f]->Observations() > 0;      // :82-84
    std::vector<int32_t> fp(N), pf(nMP);
    int nmatches = 0;
    check(orbm_search_by_projection_frame(reinterpret_cast<const orb_keypoint_t*>(&F.mvKeysUn[0]), F.mDescriptors.ptr(0), &F.mvuRight[0],
                                          &occ[0], N, &grid.cell_ptr[0], &grid.cell_idx[0], grid.bounds, &F.mvScaleFactors[0],
                                          (int)F.mvScaleFactors.size(), &mp[0], &desc[0], nMP, th, nnratio, th_high, &fp[0], &pf[0],
                                          &nmatches, device), "orbm_search_by_projection_frame");
    for(int f = 0; f < N; f++)
        if(fp[f] >= 0) F.mvpMapPoints[f] = vpMapPoints[fp[f]];                 // :121
    return nmatches;
}

// int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)  :1328-1470
// The projection statements (:1339-1378) stay as written in the reference; they fill one record per LastFrame feature:
//     orbm_proj_point_t& p = pts[i];  p.valid = 0;
//     if(pMP && !LastFrame.mvbOutlier[i]) { ... x3Dc = Rcw*x3Dw+tcw; ... if(invzc<0) continue; u = ...; v = ...;
//         if(u<mnMinX || ...) continue;
//         p.u = u; p.v = v; p.ur = u - CurrentFrame.mbf*invzc; p.octave = LastFrame.mvKeys[i].octave;
//         p.angle = LastFrame.mvKeysUn[i].angle; p.obs_positive = pMP->Observations() > 0; p.valid = 1; }
// and the window loop + rotation check become this call.  direction: 0 neither, 1 bForward, 2 bBackward (:1348-1349).
template <class Frame, class MapPoint>
int SearchByProjectionLastFrame(Frame& CurrentFrame, const FrameGrid& grid, const std::vector<orbm_proj_point_t>& pts,
                                const std::vector<MapPoint*>& lastMapPoints, float th, int direction, bool checkOri,
                                int th_high = 100, int device = 0)
{
    const int N = (int)CurrentFrame.mvKeysUn.size(), nP = (int)pts.size();
    if(N == 0 || nP == 0) return 0;
    std::vector<unsigned char> desc((size_t)nP * 32, 0), occ(N);
    for(int i = 0; i < nP; i++)
        if(pts[i].valid) std::memcpy(&desc[(size_t)i * 32], lastMapPoints[i]->GetDescriptor().ptr(0), 32);
    for(int f = 0; f < N; f++) occ[f] = CurrentFrame.mvpMapPoints[f] && CurrentFrame.mvpMapPoints[f]->Observations() > 0;   // :1404-1406
    std::vector<int32_t> fp(N), pf(nP);
    int nmatches = 0;
    check(orbm_search_by_projection_last_frame(reinterpret_cast<const orb_keypoint_t*>(&CurrentFrame.mvKeysUn[0]),
                                               CurrentFrame.mDescriptors.ptr(0), &CurrentFrame.mvuRight[0], &occ[0], N, &grid.cell_ptr[0],
                                               &grid.cell_idx[0], grid.bounds, &CurrentFrame.mvScaleFactors[0],
                                               (int)CurrentFrame.mvScaleFactors.size(), &pts[0], &desc[0], nP, th, direction, checkOri,
                                               th_high, &fp[0], &pf[0], &nmatches, device), "orbm_search_by_projection_last_frame");
    for(int f = 0; f < N; f++)
    {
        if(fp[f] >= 0) CurrentFrame.mvpMapPoints[f] = lastMapPoints[fp[f]];              // :1430
        else if(fp[f] == -2) CurrentFrame.mvpMapPoints[f] = static_cast<MapPoint*>(NULL); // :1459
    }
    return nmatches;
}

// void Frame::ComputeBoW() :400-407 / KeyFrame::ComputeBoW() KeyFrame.cc:60-69.  `voc` = orbv_create() of the loaded
// ORBVocabulary; BowVector / FeatureVector are the DBoW2 std::map types of the frame.
template <class Frame>
void ComputeBoW(Frame& F, orbv_handle_t voc, int levelsup = 4)
{
    if(!F.mBowVec.empty()) return;
    const int N = F.mDescriptors.rows;
    if(N == 0) return;
    std::vector<int32_t> word(N), node(N), bw(N), fn(N), fp(N + 1), fi(N);
    std::vector<double> weight(N), bv(N);
    int nw = 0, nn = 0;
    check(orbv_transform(voc, F.mDescriptors.ptr(0), N, levelsup, &word[0], &node[0], &weight[0]), "orbv_transform");
    check(orbv_bow_vectors(&word[0], &node[0], &weight[0], N, 1, &bw[0], &bv[0], &nw, &fn[0], &fp[0], &fi[0], &nn), "orbv_bow_vectors");
    for(int k = 0; k < nw; k++) F.mBowVec.insert(F.mBowVec.end(), std::make_pair((unsigned int)bw[k], bv[k]));
    for(int m = 0; m < nn; m++)
    {
        std::vector<unsigned int> feats(fi.begin() + fp[m], fi.begin() + fp[m + 1]);
        F.mFeatVec.insert(F.mFeatVec.end(), std::make_pair((unsigned int)fn[m], feats));
    }
}

} // namespace orbaccel

#endif
