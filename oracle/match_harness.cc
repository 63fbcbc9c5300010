// match_harness.cc -- TEST INFRASTRUCTURE ONLY.  C entry points (mh_*) that build the REFERENCE's own Frame / KeyFrame /
// MapPoint / Map objects (R21/src/{Frame,KeyFrame,MapPoint,Map}.cc compiled verbatim from /root/reference over
// oracle/cvshim + oracle/refstub) from POD arrays and run ORB_SLAM2::ORBmatcher / Frame methods on them.
//
// The file is compiled into three libraries by oracle/Makefile (same exports, different ORBmatcher / ORBextractor):
//   _ref/libmatchref.so        ORBmatcher.cc + ORBextractor.cc of the REFERENCE                      -> the ground truth
//   _ref/libmatchshim_cpu.so   cooperative-orb-slam_b200/shim/ORBmatcher.cc over oracle/orbm_on_oracle.cc -> the oracle
//                              restatement behind the product's own C++ drop-in class (pins the restatement)
//   _ref/libmatchshim_cuda.so  shim/ORBmatcher.cc + shim/ORBextractor.cc linked against liborbcuda.so -> the product
// tests/test_matcher_ref.py drives the same scenario through two of them and compares every output.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <thread>
#include <vector>
#include <list>
#include <string>
#include <cmath>
#include <algorithm>
#include <chrono>

#include <opencv2/opencv.hpp>

// the reference keeps the frame-side steps private (UndistortKeyPoints, ComputeImageBounds, AssignFeaturesToGrid) and
// calls them from its image constructors only; the harness also builds frames from given key points
#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "Map.h"
#undef private
#undef protected
#include "ORBmatcher.h"
#include "Converter.h"
#include "KeyFrameDatabase.h"

using namespace ORB_SLAM2;

// ---- the two symbols of files that are not compiled (Converter.cc needs g2o/Eigen, KeyFrameDatabase.cc DBoW2) ----
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors) {
    std::vector<cv::Mat> v;
    v.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}
void KeyFrameDatabase::erase(KeyFrame*) {}

namespace {

struct World {
    Map map;
    ORBextractor* ext_l = nullptr;
    ORBextractor* ext_r = nullptr;
    cv::Mat K, D;
    float bf = 0, th_depth = 0;
    int cols = 0, rows = 0;
    std::vector<Frame*> frames;
    // MapPoint::mObservations is a std::map<KeyFrame*, size_t>: ComputeDistinctiveDescriptors (first minimum wins) and
    // UpdateNormalAndDepth (float sum) walk it in POINTER order, so the reference's result depends on where the allocator
    // put the key frames.  Canonical rule here: key frames live in one arena, later created = larger address.
    char* kf_arena = nullptr; size_t kf_used = 0, kf_cap = 0;
    std::vector<KeyFrame*> kfs;
    std::vector<MapPoint*> mps;
    std::map<MapPoint*, int> mp_index;
    int index_of(MapPoint* p) const {
        if (!p) return -1;
        std::map<MapPoint*, int>::const_iterator it = mp_index.find(p);
        return it == mp_index.end() ? -3 : it->second;
    }
};

cv::Mat mat4(const float* T) {
    cv::Mat m(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) m.at<float>(r, c) = T[4 * r + c];
    return m;
}

}  // namespace

extern "C" {

void* mh_world_create(const float* K4, const float* dist, int ndist, float bf, float th_depth, int cols, int rows, int nfeatures,
                      float scale, int nlevels, int ini_th, int min_th) {
    World* w = new World;
    w->K = cv::Mat::eye(3, 3, CV_32F);
    w->K.at<float>(0, 0) = K4[0]; w->K.at<float>(1, 1) = K4[1]; w->K.at<float>(0, 2) = K4[2]; w->K.at<float>(1, 2) = K4[3];
    w->D = cv::Mat(ndist, 1, CV_32F);
    for (int i = 0; i < ndist; i++) w->D.at<float>(i) = dist[i];
    w->bf = bf; w->th_depth = th_depth; w->cols = cols; w->rows = rows;
    w->ext_l = new ORBextractor(nfeatures, scale, nlevels, ini_th, min_th);
    w->ext_r = new ORBextractor(nfeatures, scale, nlevels, ini_th, min_th);
    Frame::mbInitialComputations = true;      // a new calibration: the next frame recomputes the static bounds / grid cell size
    return w;
}

void mh_world_destroy(void* h) {
    World* w = (World*)h;
    if (!w) return;
    for (Frame* f : w->frames) delete f;
    // key frames / map points reference each other and the map; they are small and the process is short-lived
    delete w->ext_l; delete w->ext_r;
    delete w;
}

// Frame(imGray, ...) (R21/src/Frame.cc:176-233): extraction, undistortion, bounds, grid -- everything by the reference
int mh_frame_from_image(void* h, const uint8_t* img, size_t stride) {
    World* w = (World*)h;
    cv::Mat im(w->rows, w->cols, CV_8UC1, (void*)img, stride);
    w->frames.push_back(new Frame(im, 0.0, w->ext_l, (ORBVocabulary*)nullptr, w->K, w->D, w->bf, w->th_depth));
    return (int)w->frames.size() - 1;
}

// Frame(imLeft, imRight, ...) (:63-119): two extractor threads, undistortion, ComputeStereoMatches, grid.  Frame::mb is read
// by ComputeStereoMatches (:501) before the constructor assigns it (:116) and no constructor initialises it: the value it
// holds is whatever the storage held.  The harness decides that value (mb_prefill = the intended mbf/fx).
int mh_frame_from_stereo(void* h, const uint8_t* left, const uint8_t* right, size_t stride, float mb_prefill) {
    World* w = (World*)h;
    cv::Mat il(w->rows, w->cols, CV_8UC1, (void*)left, stride), ir(w->rows, w->cols, CV_8UC1, (void*)right, stride);
    void* mem = ::operator new(sizeof(Frame));
    memset(mem, 0, sizeof(Frame));
    memcpy((char*)mem + offsetof(Frame, mb), &mb_prefill, sizeof(float));
    Frame* f = new (mem) Frame(il, ir, 0.0, w->ext_l, w->ext_r, (ORBVocabulary*)nullptr, w->K, w->D, w->bf, w->th_depth);
    w->frames.push_back(f);
    return (int)w->frames.size() - 1;
}

// The same constructor body as :176-233 on key points / descriptors supplied by the caller instead of ExtractORB
// (u_right / depth: NULL for a monocular frame).  UndistortKeyPoints, ComputeImageBounds and AssignFeaturesToGrid are the
// reference's own methods.
int mh_frame_from_features(void* h, const orc_keypoint* kps, const uint8_t* desc, int n, const float* u_right, const float* depth) {
    World* w = (World*)h;
    Frame* f = new Frame();
    f->mpORBvocabulary = nullptr; f->mpORBextractorLeft = w->ext_l; f->mpORBextractorRight = nullptr;
    f->mTimeStamp = 0.0; f->mK = w->K.clone(); f->mDistCoef = w->D.clone(); f->mbf = w->bf; f->mThDepth = w->th_depth;
    f->mnId = Frame::nNextId++;
    f->mnScaleLevels = w->ext_l->GetLevels();
    f->mfScaleFactor = w->ext_l->GetScaleFactor();
    f->mfLogScaleFactor = log(f->mfScaleFactor);
    f->mvScaleFactors = w->ext_l->GetScaleFactors();
    f->mvInvScaleFactors = w->ext_l->GetInverseScaleFactors();
    f->mvLevelSigma2 = w->ext_l->GetScaleSigmaSquares();
    f->mvInvLevelSigma2 = w->ext_l->GetInverseScaleSigmaSquares();
    f->mvKeys.resize(n);
    if (n) memcpy(&f->mvKeys[0], kps, (size_t)n * sizeof(cv::KeyPoint));
    f->mDescriptors.create(n, 32, CV_8U);
    if (n) memcpy(f->mDescriptors.ptr(0), desc, (size_t)n * 32);
    f->N = n;
    w->frames.push_back(f);
    if (n == 0) return (int)w->frames.size() - 1;
    f->UndistortKeyPoints();
    f->mvuRight = u_right ? std::vector<float>(u_right, u_right + n) : std::vector<float>(n, -1);
    f->mvDepth = depth ? std::vector<float>(depth, depth + n) : std::vector<float>(n, -1);
    f->mvpMapPoints = std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
    f->mvbOutlier = std::vector<bool>(n, false);
    if (Frame::mbInitialComputations) {
        cv::Mat im(w->rows, w->cols, CV_8UC1);
        f->ComputeImageBounds(im);
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        Frame::fx = w->K.at<float>(0, 0); Frame::fy = w->K.at<float>(1, 1);
        Frame::cx = w->K.at<float>(0, 2); Frame::cy = w->K.at<float>(1, 2);
        Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
        Frame::mbInitialComputations = false;
    }
    f->mb = f->mbf / Frame::fx;
    f->AssignFeaturesToGrid();
    return (int)w->frames.size() - 1;
}

int mh_frame_n(void* h, int fi) { return ((World*)h)->frames[fi]->N; }
int mh_frame_n_right(void* h, int fi) { return (int)((World*)h)->frames[fi]->mvKeysRight.size(); }

// any pointer may be NULL.  bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY); cell_ptr [64*48+1], cell_idx [N] (cell = ix*48 + iy)
void mh_frame_get(void* h, int fi, orc_keypoint* kps, orc_keypoint* kps_un, uint8_t* desc, float* u_right, float* depth, float* bounds,
                  int32_t* cell_ptr, int32_t* cell_idx, orc_keypoint* kps_right, uint8_t* desc_right) {
    Frame* f = ((World*)h)->frames[fi];
    const int n = f->N;
    if (kps && n) memcpy(kps, &f->mvKeys[0], (size_t)n * sizeof(cv::KeyPoint));
    if (kps_un && n) memcpy(kps_un, &f->mvKeysUn[0], (size_t)n * sizeof(cv::KeyPoint));
    if (desc) for (int i = 0; i < n; i++) memcpy(desc + (size_t)i * 32, f->mDescriptors.ptr(i), 32);
    if (u_right && n) memcpy(u_right, &f->mvuRight[0], (size_t)n * 4);
    if (depth && n) memcpy(depth, &f->mvDepth[0], (size_t)n * 4);
    if (bounds) { bounds[0] = Frame::mnMinX; bounds[1] = Frame::mnMaxX; bounds[2] = Frame::mnMinY; bounds[3] = Frame::mnMaxY; }
    if (cell_ptr && cell_idx) {
        int at = 0, c = 0;
        for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
            for (int iy = 0; iy < FRAME_GRID_ROWS; iy++, c++) {
                cell_ptr[c] = at;
                for (size_t k = 0; k < f->mGrid[ix][iy].size(); k++) cell_idx[at++] = (int32_t)f->mGrid[ix][iy][k];
            }
        cell_ptr[c] = at;
    }
    const int nr = (int)f->mvKeysRight.size();
    if (kps_right && nr) memcpy(kps_right, &f->mvKeysRight[0], (size_t)nr * sizeof(cv::KeyPoint));
    if (desc_right) for (int i = 0; i < nr; i++) memcpy(desc_right + (size_t)i * 32, f->mDescriptorsRight.ptr(i), 32);
}

void mh_frame_set_pose(void* h, int fi, const float* Tcw) { ((World*)h)->frames[fi]->SetPose(mat4(Tcw)); }

static void fill_featvec(DBoW2::FeatureVector& fv, int n_nodes, const int32_t* ids, const int32_t* ptr, const int32_t* idx) {
    fv.clear();
    for (int k = 0; k < n_nodes; k++)
        for (int j = ptr[k]; j < ptr[k + 1]; j++) fv.addFeature((DBoW2::NodeId)ids[k], (unsigned int)idx[j]);
}
void mh_frame_set_featvec(void* h, int fi, int n_nodes, const int32_t* ids, const int32_t* ptr, const int32_t* idx) {
    fill_featvec(((World*)h)->frames[fi]->mFeatVec, n_nodes, ids, ptr, idx);
}
// mvpMapPoints[idx] = map point mp (-1: NULL), mvbOutlier[idx] = outlier
void mh_frame_set_mappoint(void* h, int fi, int idx, int mp, int outlier) {
    World* w = (World*)h;
    w->frames[fi]->mvpMapPoints[idx] = mp < 0 ? static_cast<MapPoint*>(NULL) : w->mps[mp];
    w->frames[fi]->mvbOutlier[idx] = outlier != 0;
}
void mh_frame_get_mappoints(void* h, int fi, int32_t* out) {
    World* w = (World*)h;
    Frame* f = w->frames[fi];
    for (int i = 0; i < f->N; i++) out[i] = w->index_of(f->mvpMapPoints[i]);
}
int mh_frame_features_in_area(void* h, int fi, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    const std::vector<size_t> v = ((World*)h)->frames[fi]->GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// KeyFrame(F, pMap, pKFDB) (R21/src/KeyFrame.cc:31-58): copies the frame (pose, key points, grid, feature vector)
int mh_keyframe(void* h, int fi) {
    World* w = (World*)h;
    const size_t slot = (sizeof(KeyFrame) + 63) & ~(size_t)63;
    if (!w->kf_arena) { w->kf_cap = 64 * slot; w->kf_arena = (char*)aligned_alloc(64, w->kf_cap); }
    if (w->kf_used + slot > w->kf_cap) return -1;
    KeyFrame* kf = new (w->kf_arena + w->kf_used) KeyFrame(*w->frames[fi], &w->map, (KeyFrameDatabase*)nullptr);
    w->kf_used += slot;
    w->map.AddKeyFrame(kf);
    w->kfs.push_back(kf);
    return (int)w->kfs.size() - 1;
}
int mh_keyframe_n(void* h, int ki) { return ((World*)h)->kfs[ki]->N; }
void mh_keyframe_get_mappoints(void* h, int ki, int32_t* out) {
    World* w = (World*)h;
    const std::vector<MapPoint*> v = w->kfs[ki]->GetMapPointMatches();
    for (size_t i = 0; i < v.size(); i++) out[i] = w->index_of(v[i]);
}
int mh_keyframe_features_in_area(void* h, int ki, float x, float y, float r, int32_t* out, int cap) {
    const std::vector<size_t> v = ((World*)h)->kfs[ki]->GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}
void mh_keyframe_bounds(void* h, int ki, int32_t* out4) {
    KeyFrame* kf = ((World*)h)->kfs[ki];
    out4[0] = kf->mnMinX; out4[1] = kf->mnMaxX; out4[2] = kf->mnMinY; out4[3] = kf->mnMaxY;
}

// MapPoint(Pos, pRefKF, pMap) (R21/src/MapPoint.cc:32-45)
int mh_mappoint(void* h, const float* pos, int ref_kf) {
    World* w = (World*)h;
    cv::Mat p(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) p.at<float>(i) = pos[i];
    MapPoint* mp = new MapPoint(p, w->kfs[ref_kf], &w->map);
    w->map.AddMapPoint(mp);
    w->mp_index[mp] = (int)w->mps.size();
    w->mps.push_back(mp);
    return (int)w->mps.size() - 1;
}
// the pair of calls LocalMapping makes for a new observation (LocalMapping.cc:430-433)
void mh_observe(void* h, int mp, int ki, int idx) {
    World* w = (World*)h;
    w->mps[mp]->AddObservation(w->kfs[ki], (size_t)idx);
    w->kfs[ki]->AddMapPoint(w->mps[mp], (size_t)idx);
}
// ComputeDistinctiveDescriptors (MapPoint.cc:242-307) + UpdateNormalAndDepth (:332-377)
void mh_mappoint_update(void* h, int mp) {
    MapPoint* p = ((World*)h)->mps[mp];
    p->ComputeDistinctiveDescriptors();
    p->UpdateNormalAndDepth();
}
// desc [32], normal [3], dist [2] = (min, max distance invariance), state [3] = (observations, isBad, index of GetReplaced())
void mh_mappoint_get(void* h, int mp, uint8_t* desc, float* normal, float* dist, int32_t* state) {
    World* w = (World*)h;
    MapPoint* p = w->mps[mp];
    if (desc) { cv::Mat d = p->GetDescriptor(); if (!d.empty()) memcpy(desc, d.ptr(0), 32); else memset(desc, 0, 32); }
    if (normal) { cv::Mat n = p->GetNormal(); for (int i = 0; i < 3; i++) normal[i] = n.at<float>(i); }
    if (dist) { dist[0] = p->GetMinDistanceInvariance(); dist[1] = p->GetMaxDistanceInvariance(); }
    if (state) { state[0] = p->Observations(); state[1] = p->isBad() ? 1 : 0; state[2] = w->index_of(p->GetReplaced()); }
}
void mh_mappoint_set_bad(void* h, int mp) { ((World*)h)->mps[mp]->SetBadFlag(); }

// ---- the matcher ---------------------------------------------------------------------------------------------------
int mh_search_by_bow_kf_f(void* h, int ki, int fi, float nnratio, int check_ori, int32_t* out) {
    World* w = (World*)h;
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint*> res;
    const int n = m.SearchByBoW(w->kfs[ki], *w->frames[fi], res);
    for (size_t i = 0; i < res.size(); i++) out[i] = w->index_of(res[i]);
    return n;
}
int mh_search_by_bow_kf_kf(void* h, int k1, int k2, float nnratio, int check_ori, int32_t* out) {
    World* w = (World*)h;
    ORBmatcher m(nnratio, check_ori != 0);
    std::vector<MapPoint*> res;
    const int n = m.SearchByBoW(w->kfs[k1], w->kfs[k2], res);
    for (size_t i = 0; i < res.size(); i++) out[i] = w->index_of(res[i]);
    return n;
}
int mh_search_for_triangulation(void* h, int k1, int k2, const float* F12, int only_stereo, float nnratio, int check_ori, int32_t* pairs,
                                int cap) {
    World* w = (World*)h;
    ORBmatcher m(nnratio, check_ori != 0);
    cv::Mat F(3, 3, CV_32F);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F.at<float>(r, c) = F12[3 * r + c];
    std::vector<std::pair<size_t, size_t> > v;
    const int n = m.SearchForTriangulation(w->kfs[k1], w->kfs[k2], F, v, only_stereo != 0);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) { pairs[2 * i] = (int32_t)v[i].first; pairs[2 * i + 1] = (int32_t)v[i].second; }
    return n;
}
// Tracking::SearchLocalPoints (Tracking.cc:1143-1192): isInFrustum marks the points, then SearchByProjection(F, points, th)
int mh_search_by_projection_local(void* h, int fi, const int32_t* mp_list, int n, float th, float nnratio, int32_t* out, int32_t* in_view) {
    World* w = (World*)h;
    Frame& F = *w->frames[fi];
    std::vector<MapPoint*> v(n);
    for (int i = 0; i < n; i++) {
        v[i] = w->mps[mp_list[i]];
        v[i]->mbTrackInView = false;
        const bool iv = F.isInFrustum(v[i], 0.5);
        if (in_view) in_view[i] = iv ? 1 : 0;
    }
    ORBmatcher m(nnratio, true);
    const int r = m.SearchByProjection(F, v, th);
    for (int i = 0; i < F.N; i++) out[i] = w->index_of(F.mvpMapPoints[i]);
    return r;
}
int mh_search_by_projection_last(void* h, int cur, int last, float th, int mono, float nnratio, int check_ori, int32_t* out) {
    World* w = (World*)h;
    Frame& F = *w->frames[cur];
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchByProjection(F, *w->frames[last], th, mono != 0);
    for (int i = 0; i < F.N; i++) out[i] = w->index_of(F.mvpMapPoints[i]);
    return r;
}
int mh_search_by_projection_kf(void* h, int cur, int ki, const int32_t* already, int n_already, float th, int orb_dist, float nnratio,
                               int check_ori, int32_t* out) {
    World* w = (World*)h;
    Frame& F = *w->frames[cur];
    std::set<MapPoint*> found;
    for (int i = 0; i < n_already; i++) found.insert(w->mps[already[i]]);
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchByProjection(F, w->kfs[ki], found, th, orb_dist);
    for (int i = 0; i < F.N; i++) out[i] = w->index_of(F.mvpMapPoints[i]);
    return r;
}
// matched [KF.N]: map point indices (-1: NULL) in, updated out
int mh_search_by_projection_sim3(void* h, int ki, const float* Scw, const int32_t* points, int np, int32_t* matched, int th) {
    World* w = (World*)h;
    KeyFrame* kf = w->kfs[ki];
    std::vector<MapPoint*> pts(np), m(kf->N);
    for (int i = 0; i < np; i++) pts[i] = w->mps[points[i]];
    for (int i = 0; i < kf->N; i++) m[i] = matched[i] < 0 ? static_cast<MapPoint*>(NULL) : w->mps[matched[i]];
    ORBmatcher matcher(0.75f, true);
    const int r = matcher.SearchByProjection(kf, mat4(Scw), pts, m, th);
    for (int i = 0; i < kf->N; i++) matched[i] = w->index_of(m[i]);
    return r;
}
int mh_search_for_initialization(void* h, int f1, int f2, float* prev_xy, int window, float nnratio, int check_ori, int32_t* m12) {
    World* w = (World*)h;
    Frame& F1 = *w->frames[f1];
    std::vector<cv::Point2f> prev(F1.mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> v;
    ORBmatcher m(nnratio, check_ori != 0);
    const int r = m.SearchForInitialization(F1, *w->frames[f2], prev, v, window);
    for (size_t i = 0; i < v.size(); i++) m12[i] = v[i];
    for (size_t i = 0; i < prev.size(); i++) { prev_xy[2 * i] = prev[i].x; prev_xy[2 * i + 1] = prev[i].y; }
    return r;
}
// Fuse(pKF, vpMapPoints, th) (:825-975); mp_list entries < 0 are NULL pointers.  The resulting map state is read back with
// mh_keyframe_get_mappoints / mh_mappoint_get.
int mh_fuse(void* h, int ki, const int32_t* mp_list, int n, float th) {
    World* w = (World*)h;
    std::vector<MapPoint*> v(n);
    for (int i = 0; i < n; i++) v[i] = mp_list[i] < 0 ? static_cast<MapPoint*>(NULL) : w->mps[mp_list[i]];
    ORBmatcher m(0.6f, true);
    return m.Fuse(w->kfs[ki], v, th);
}
int mh_fuse_sim3(void* h, int ki, const float* Scw, const int32_t* mp_list, int n, float th, int32_t* replace) {
    World* w = (World*)h;
    std::vector<MapPoint*> v(n), rep(n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n; i++) v[i] = w->mps[mp_list[i]];
    ORBmatcher m(0.8f, true);
    const int r = m.Fuse(w->kfs[ki], mat4(Scw), v, th, rep);
    for (int i = 0; i < n; i++) replace[i] = w->index_of(rep[i]);
    return r;
}
// matches12 [KF1.N]: map point indices of KF2 (-1: NULL) in, updated out
int mh_search_by_sim3(void* h, int k1, int k2, int32_t* matches12, float s12, const float* R12, const float* t12, float th) {
    World* w = (World*)h;
    KeyFrame* kf1 = w->kfs[k1];
    std::vector<MapPoint*> m12(kf1->N);
    for (int i = 0; i < kf1->N; i++) m12[i] = matches12[i] < 0 ? static_cast<MapPoint*>(NULL) : w->mps[matches12[i]];
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    for (int r = 0; r < 3; r++) { t.at<float>(r) = t12[r]; for (int c = 0; c < 3; c++) R.at<float>(r, c) = R12[3 * r + c]; }
    ORBmatcher m(0.75f, true);
    const int n = m.SearchBySim3(kf1, w->kfs[k2], m12, s12, R, t, th);
    for (int i = 0; i < kf1->N; i++) matches12[i] = w->index_of(m12[i]);
    return n;
}
int mh_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat ma(1, 32, CV_8U, (void*)a), mb(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

// per-frame latency of ORBextractor::operator() as Tracking calls it (Tracking.cc:258-260 -> Frame.cc:252-258), wall clock, in
// milliseconds: mirror = 0 / 1 switches the mvImagePyramid host mirror of the drop-in (ignored by the reference build)
double mh_extract_latency_ms(void* h, const uint8_t* img, size_t stride, int iters, int mirror, int* n_keys) {
    World* w = (World*)h;
    cv::Mat im(w->rows, w->cols, CV_8UC1, (void*)img, stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
#ifdef MH_CUDA_DROPIN
    w->ext_l->SetPyramidMirror(mirror != 0);
#else
    (void)mirror;
#endif
    for (int i = 0; i < 3; i++) (*w->ext_l)(im, cv::Mat(), keys, desc);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) (*w->ext_l)(im, cv::Mat(), keys, desc);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    if (n_keys) *n_keys = (int)keys.size();
#ifdef MH_CUDA_DROPIN
    w->ext_l->SetPyramidMirror(true);
#endif
    return ms;
}

#ifdef MH_CUDA_DROPIN
// The product's C++ drop-in only: Frame::ComputeStereoMatches routed to the GPU (orbaccel::ComputeStereoMatches reads the
// pyramids on the device), beside the reference's own ComputeStereoMatches that ran inside the stereo constructor on the
// extractor's host mirror.  Returns the match count; u_right / depth [N].
}  // extern "C"
#include "ORBmatcher_accel.h"
extern "C" {
int mh_frame_stereo_accel(void* h, int fi, float* u_right, float* depth) {
    Frame F(*((World*)h)->frames[fi]);
    orbaccel::ComputeStereoMatches(F);
    int n = 0;
    for (int i = 0; i < F.N; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; n += F.mvuRight[i] >= 0; }
    return n;
}
#endif

}  // extern "C"
