/*
 * orbcuda.h -- C ABI of liborbcuda.so: the B200 (sm_100a) ORB front-end.
 *
 * Drop-in boundary for the reference's ORBextractor / ORBmatcher hot path.  The reference
 * (R21 = ORB_SLAM2.1 in 530300865/Cooperative-ORB-SLAM) has no FFI layer: the boundary is its C++
 * class API, so each entry point below names the reference interface it replaces (file:line) and
 * the C++ shim classes in cooperative-orb-slam_b200/shim/ forward to it 1:1.  POD only: plain
 * pointers and sizes, no cv:: or torch types, no exceptions.  Every function returns an int status
 * (ORB_OK == 0); there is NO CPU fallback -- without a CUDA device every compute call returns
 * ORB_ERR_CUDA.
 *
 * Threading: one orbx handle per ORBextractor instance; calls on one handle must be serialised by
 * the caller, different handles may be used concurrently (the reference runs left/right extractors
 * on two threads, R21/src/Frame.cc:80-83).  Matcher entry points are re-entrant.
 */
#ifndef ORBCUDA_H
#define ORBCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_OK 0
#define ORB_ERR_ARG 1        /* bad argument / unsupported shape */
#define ORB_ERR_CUDA 2       /* CUDA runtime error or no device  */
#define ORB_ERR_CAPACITY 3   /* caller buffer too small          */
#define ORB_MAX_LEVELS 16

/* Same 28-byte layout as cv::KeyPoint {Point2f pt; float size, angle, response; int octave, class_id}
 * so a shim can memcpy into std::vector<cv::KeyPoint>. */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orb_keypoint_t;

/* ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
 * R21/src/ORBextractor.cc:410-470 */
typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
} orbx_params_t;

typedef struct orbx_handle_s* orbx_handle_t;

const char* orb_last_error(void);          /* thread-local text of the last failure */
int orb_device_count(int* count);

/* ---------------------------------------------------------------- extractor ------------------ */
/* Creates an extractor bound to CUDA device `device` for images up to max_width x max_height and up
 * to max_batch frames per call.  Owns device buffers, pinned staging and one stream. */
int orbx_create(const orbx_params_t* params, int max_width, int max_height, int max_batch, int device,
                orbx_handle_t* out);
int orbx_destroy(orbx_handle_t h);

/* GetLevels/GetScaleFactor(s)/GetInverseScaleFactors/GetScaleSigmaSquares/GetInverseScaleSigmaSquares
 * (R21/include/ORBextractor.h:63-83) plus mnFeaturesPerLevel (:435-446).  Any pointer may be NULL.
 * Pure host code (no CUDA call). */
int orbx_tables(orbx_handle_t h, float* scale_factors, float* inv_scale_factors, float* level_sigma2,
                float* inv_level_sigma2, int32_t* features_per_level);
/* Upper bound of keypoints one frame can produce (sum over levels of max(N_l + 3, 4*nIni)). */
int orbx_max_keypoints(orbx_handle_t h, int width, int height, int* cap);

/* ORBextractor::operator()(image, mask, keypoints, descriptors)  R21/src/ORBextractor.cc:1043-1105.
 * image: CV_8UC1, `stride` bytes per row, host memory.  The mask is ignored by the reference and has
 * no parameter here.  keypoints[cap] / descriptors[cap*32] are caller-owned host buffers;
 * *n_keypoints receives the count (level-major, quadtree list order inside a level).  An empty image
 * (NULL / 0 size) returns ORB_OK with 0 keypoints like :1046-1047.  Synchronous. */
int orbx_extract(orbx_handle_t h, const uint8_t* image, int width, int height, size_t stride,
                 orb_keypoint_t* keypoints, uint8_t* descriptors, int cap, int* n_keypoints);

/* Batched form for agent streams: n_frames images of identical size; frame i is at
 * images + i*frame_stride.  Outputs are [n_frames][cap] / [n_frames][cap][32] / [n_frames].
 * orbx_extract_batch_async enqueues upload + kernels + download on the handle's stream and returns;
 * orbx_wait blocks until the results are in the caller's buffers (which must be pinned or stay
 * alive until then).  orbx_extract_batch == async + wait. */
int orbx_extract_batch(orbx_handle_t h, const uint8_t* images, int n_frames, int width, int height,
                       size_t row_stride, size_t frame_stride, orb_keypoint_t* keypoints,
                       uint8_t* descriptors, int cap, int32_t* n_keypoints);
int orbx_extract_batch_async(orbx_handle_t h, const uint8_t* images, int n_frames, int width, int height,
                             size_t row_stride, size_t frame_stride, orb_keypoint_t* keypoints,
                             uint8_t* descriptors, int cap, int32_t* n_keypoints);
int orbx_wait(orbx_handle_t h);

/* Device-resident form: images already in HBM (device pointer, same strides), results stay in HBM.
 * d_keypoints [n_frames][cap], d_descriptors [n_frames][cap][32], d_counts [n_frames] are device
 * pointers; enqueued on the handle's stream (orbx_wait to finish). */
int orbx_extract_batch_device(orbx_handle_t h, const uint8_t* d_images, int n_frames, int width, int height,
                              size_t row_stride, size_t frame_stride, orb_keypoint_t* d_keypoints,
                              uint8_t* d_descriptors, int cap, int32_t* d_counts);

/* Host pinned memory helpers so callers can make uploads truly asynchronous. */
int orb_host_alloc(void** ptr, size_t bytes);
int orb_host_free(void* ptr);

/* ---- views of the last call's intermediate results (public mvImagePyramid + parity tests) ---- */
/* Level geometry of the last extraction. */
int orbx_level_size(orbx_handle_t h, int level, int* width, int* height);
/* mvImagePyramid[level] (R21/include/ORBextractor.h:85; ROI of the padded plane :1113-1116).
 * with_border != 0 copies the (w+38) x (h+38) padded plane, else the w x h ROI. */
int orbx_download_level(orbx_handle_t h, int frame, int level, int with_border, uint8_t* dst, size_t dst_stride);
/* The same for every level of one frame at once (the whole mvImagePyramid mirror): dst[l] / dst_stride[l] = plane and row stride
 * of level l, (w_l + 38) x (h_l + 38) with the border.  One stream synchronisation instead of one per level; the copies are
 * asynchronous when the planes are page-locked (orb_host_alloc). */
int orbx_download_pyramid(orbx_handle_t h, int frame, int with_border, uint8_t* const* dst, const size_t* dst_stride);
/* The 7x7 sigma=2 blurred level used for descriptors (R21 :1085-1086). */
int orbx_download_blurred(orbx_handle_t h, int frame, int level, uint8_t* dst, size_t dst_stride);
/* vToDistributeKeys of a level (R21 :789-826): candidates in cell-major, raster-in-cell order;
 * coordinates relative to (minBorderX, minBorderY) = (16,16), response = FAST score. */
int orbx_download_candidates(orbx_handle_t h, int frame, int level, int16_t* x, int16_t* y, uint8_t* score,
                             int cap, int* n);
/* The per-pixel FAST score map at minThFAST (0 where not a corner), w x h of the level. */
int orbx_download_scores(orbx_handle_t h, int frame, int level, uint8_t* dst, size_t dst_stride);

/* CUDA-event stage timers of the last batch, milliseconds:
 * [0]=upload [1]=pyramid [2]=fast score [3]=blur [4]=cell nms [5]=quadtree [6]=orient+describe
 * [7]=download [8]=total.  Enabled with orbx_set_profiling(h,1) (adds event records). */
int orbx_set_profiling(orbx_handle_t h, int enable);
int orbx_stage_times(orbx_handle_t h, float* ms9);
/* Number of kernel launches issued by this handle since creation. */
int orbx_launch_count(orbx_handle_t h, int64_t* launches);
/* The CUDA stream of the handle as an opaque pointer (cudaStream_t). */
int orbx_stream(orbx_handle_t h, void** stream);

/* ---- stand-alone stage entry points (tests, stage benchmarks); host in/out, synchronous ---- */
/* DistributeOctTree (R21 :539-763) on caller-provided candidates (coords relative to min_x/min_y). */
int orbx_distribute_octtree(const int16_t* x, const int16_t* y, const uint8_t* score, int n, int min_x,
                            int max_x, int min_y, int max_y, int n_features, int32_t* out_index, int cap,
                            int* n_out, int device);

/* ---------------------------------------------------------------- matcher -------------------- */
/* ORBmatcher::DescriptorDistance(a, b)  R21/src/ORBmatcher.cc:1647-1663 (host, bit-exact). */
int orb_hamming256(const void* a, const void* b);

/* Brute-force 2-NN over 256-bit descriptors with the reference's best/second-best rule
 * (R21/src/ORBmatcher.cc:216-225: strict '<', both start at 256, first index wins ties).
 * queries [nq][32], map [nm][32] host memory.  Outputs per query: best index (+index_base, -1 if
 * nm==0), best distance, second-best distance, and (optional, may be NULL) the index of the second
 * best in (distance, index) lexicographic order.  variant selects the kernel (identical results):
 *   0  LOP3+POPC, one query per thread (XU/POPC-pipe bound, ~0.5 Tcmp/s on B200)
 *   1  tensor-core AND-popc contraction d = popc(a)+popc(b)-2*popc(a&b) on mma.sync integer MMAs (~1.0 Tcmp/s)
 *   2  same, streaming the map without shared memory (~0.8 Tcmp/s; kept as evidence)
 *   3  the contraction on tcgen05.mma kind::i8 with TMEM accumulators, warp-specialised mbarrier pipeline
 *      (~6.1 Tcmp/s)
 *   4  as 3 with the query operand held in TMEM (~6.1 Tcmp/s)
 *   5  as 3 on CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2: each CTA expands half of every map tile;
 *      ~6.8 Tcmp/s; the default of the Python/C++ host layers).  See DESIGN.md section 4 and profiles/. */
int orbm_knn2(const uint8_t* queries, int nq, const uint8_t* map, int64_t nm, int64_t index_base,
              int32_t* best_idx, int32_t* best_dist, int32_t* second_dist, int32_t* second_idx,
              int variant, int device);
/* Same with device pointers, asynchronous on `stream` (cudaStream_t; NULL = default stream).
 * d_out is [nq] x int32[4] = {d1, i1, d2, i2}: the per-rank record exchanged by the sharded map
 * match (allgather) and merged with orbm_merge_top2_device. */
int orbm_knn2_device(const uint8_t* d_queries, int nq, const uint8_t* d_map, int64_t nm, int64_t index_base,
                     int32_t* d_out, int variant, void* stream);
/* Merge `parts` records per query ([parts][nq][4] int32, device) in lexicographic (dist, index)
 * order into [nq][4]; bit-identical to a single-rank search over the concatenated map. */
int orbm_merge_top2_device(const int32_t* d_parts, int parts, int nq, int32_t* d_out, void* stream);
int orbm_merge_top2_host(const int32_t* parts_rec, int parts, int nq, int32_t* out);
/* Acceptance test of R21 ORBmatcher.cc:228-230 on merged records: match iff d1 <= th (or < th when
 * strict) and (float)d1 < ratio*(float)d2.  out_match[i] = i1 or -1. */
int orbm_ratio_test_host(const int32_t* rec, int nq, float ratio, int th, int strict, int32_t* out_match);
/* The same acceptance test on the device.  orbm_knn2_ratio_device = orbm_knn2_device + the test in one call chain (the
 * test is folded into the kernel that merges the map splits): d_match[q] = accepted nearest neighbour or -1; d_rec may be
 * NULL.  orbm_ratio_test_device applies it to records already on the device (e.g. a sharded search's merged records). */
int orbm_knn2_ratio_device(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, float ratio,
                           int th, int strict, int32_t* d_rec, int32_t* d_match, int variant, void* stream);
int orbm_ratio_test_device(const int32_t* d_rec, int nq, float ratio, int th, int strict, int32_t* d_match, void* stream);

/* void MapPoint::ComputeDistinctiveDescriptors()  R21/src/MapPoint.cc:242-307, batched over map points.
 * Point p owns the descriptors of its (non-bad) observations desc[ptr[p] .. ptr[p+1]) (N x 32 bytes, host);
 * best[p] = index inside that list of the descriptor with the least median Hamming distance to the others
 * (median = sorted[(int)(0.5*(N-1))], first minimum wins), -1 for a point without observations. */
int orbm_distinctive_descriptors(const uint8_t* desc, const int32_t* ptr, int n_points, int32_t* best, int device);

/* DBoW2::FeatureVector as CSR: node ids ascending; node i owns idx[ptr[i] .. ptr[i+1]). */
typedef struct {
    int32_t n_nodes;
    const int32_t* node_ids;
    const int32_t* ptr;
    const int32_t* idx;
} orbm_featvec_t;

/* int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 * R21/src/ORBmatcher.cc:159-288.  kf_valid[i] != 0 <=> KF feature i has a non-bad MapPoint.
 * out_match_f[j] = KF feature index whose MapPoint is assigned to frame feature j, or -1.
 * *n_matches = return value. */
int orbm_search_by_bow_kf_f(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid, int n_kf,
                            const orbm_featvec_t* fv_kf, const uint8_t* desc_f, const float* angle_f, int n_f,
                            const orbm_featvec_t* fv_f, float nnratio, int check_orientation,
                            int32_t* out_match_f, int* n_matches, int device);
/* int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 * R21/src/ORBmatcher.cc:522-655.  out_match12[i1] = feature index in KF2 or -1. */
int orbm_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                             const orbm_featvec_t* fv1, const uint8_t* desc2, const float* angle2,
                             const uint8_t* valid2, int n2, const orbm_featvec_t* fv2, float nnratio,
                             int check_orientation, int32_t* out_match12, int* n_matches, int device);

typedef struct {
    float x, y;       /* mvKeysUn[i].pt      */
    float angle;      /* mvKeysUn[i].angle   */
    int32_t octave;   /* mvKeysUn[i].octave  */
    float u_right;    /* mvuRight[i] (<0: monocular) */
    int32_t has_mp;   /* GetMapPoint(i) != NULL */
} orbm_tri_feature_t;

/* int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
 *       vector<pair<size_t,size_t>>& vMatchedPairs, const bool bOnlyStereo)
 * R21/src/ORBmatcher.cc:657-823.  F12 row-major 3x3 float; (ex,ey) = epipole in image 2 (:664-670);
 * scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2.  out_pairs = (idx1, idx2)
 * ascending idx1 (:812-820). */
int orbm_search_for_triangulation(const uint8_t* desc1, const orbm_tri_feature_t* f1, int n1,
                                  const orbm_featvec_t* fv1, const uint8_t* desc2,
                                  const orbm_tri_feature_t* f2, int n2, const orbm_featvec_t* fv2,
                                  const float* F12, float ex, float ey, const float* scale_factors2,
                                  const float* level_sigma2_2, int only_stereo, int check_orientation,
                                  int32_t* out_pairs, int cap_pairs, int* n_matches, int device);

/* void Frame::ComputeStereoMatches()  R21/src/Frame.cc:471-645.  hl/hr: the two extractor handles
 * whose last orbx_extract produced the left/right keypoints (their pyramids are read on the device,
 * replacing the reads of mpORBextractorLeft/Right->mvImagePyramid, :568,:585).  mb = mbf/fx is
 * passed explicitly.  u_right[n_left], depth[n_left] = mvuRight / mvDepth. */
int orbm_stereo_matches(orbx_handle_t hl, orbx_handle_t hr, const orb_keypoint_t* keys_left,
                        const uint8_t* desc_left, int n_left, const orb_keypoint_t* keys_right,
                        const uint8_t* desc_right, int n_right, float mbf, float mb, float* u_right,
                        float* depth, int* n_matches);

/* The same for a batch of stereo pairs without leaving the device (two extractors per stereo frame, R21/src/Frame.cc:80-83,
 * then ComputeStereoMatches :471-645, outlier cut :630-644 included).  hl/hr hold the pyramids of their last
 * orbx_extract_batch_device/_async call (>= n_pairs frames); d_keys_* / d_desc_* / d_counts_* are the device outputs of
 * those calls ([n_pairs][cap] layout).  d_scratch: orbm_stereo_scratch_bytes(n_pairs, cap) bytes of device memory.
 * d_u_right / d_depth: [n_pairs][cap] floats (-1 where a key point has no match), d_n_matches: [n_pairs].  Enqueued on
 * `stream`, which is made to wait for both extractors' streams; nothing is synchronised. */
int orbm_stereo_matches_batch_device(orbx_handle_t hl, orbx_handle_t hr, const void* d_keys_left,
                                     const uint8_t* d_desc_left, const int32_t* d_counts_left,
                                     const void* d_keys_right, const uint8_t* d_desc_right,
                                     const int32_t* d_counts_right, int n_pairs, int cap, float mbf, float mb,
                                     void* d_scratch, float* d_u_right, float* d_depth, int32_t* d_n_matches,
                                     void* stream);
size_t orbm_stereo_scratch_bytes(int n_pairs, int cap);

/* ---------------------------------------------------------------- frame (the steps after extraction) ---- */
#define ORBF_GRID_COLS 64   /* FRAME_GRID_COLS, R21/include/Frame.h:37 */
#define ORBF_GRID_ROWS 48   /* FRAME_GRID_ROWS, R21/include/Frame.h:36 */

/* void Frame::UndistortKeyPoints()  R21/src/Frame.cc:409-439: cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)
 * on the key point coordinates (5 fixed-point iterations in double), everything else of the key point copied; a
 * plain copy when dist[0] == 0 (:411-415).  K = (fx, fy, cx, cy) as in mK (CV_32F), dist = mDistCoef (k1, k2, p1, p2[, k3]). */
int orbf_undistort_keypoints(const orb_keypoint_t* kps, int n, const float* K, const float* dist, int ndist,
                             orb_keypoint_t* out, int device);
/* void Frame::ComputeImageBounds(const cv::Mat& imLeft)  R21/src/Frame.cc:441-470.
 * bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY). */
int orbf_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds, int device);
/* void Frame::AssignFeaturesToGrid()  R21/src/Frame.cc:235-250 with PosInGrid :387-397 (round, not floor).  The
 * grid is returned as CSR over the cells in mGrid[ix][iy] order (cell = ix * 48 + iy): cell_ptr[64*48 + 1],
 * cell_idx[n]; indices ascend inside a cell (push_back order).  *n_assigned = key points inside the grid. */
int orbf_assign_grid(const orb_keypoint_t* kps_un, int n, const float* bounds, int32_t* cell_ptr, int32_t* cell_idx,
                     int* n_assigned, int device);
/* Frame::UndistortKeyPoints (:409-439) followed by Frame::AssignFeaturesToGrid (:235-250), the two steps the Frame constructor
 * (R21/src/Frame.cc:58-116, :119-174, :177-232) runs back to back on the extractor's key points, as ONE pass over the device:
 * one upload of the key points, two kernels, one download of out_kps_un[n], cell_ptr[64*48 + 1] and cell_idx[n].  bounds = the
 * result of orbf_image_bounds for this camera (the reference computes it once: mbInitialComputations). */
int orbf_build_frame(const orb_keypoint_t* kps, int n, const float* K, const float* dist, int ndist, const float* bounds,
                     orb_keypoint_t* out_kps_un, int32_t* cell_ptr, int32_t* cell_idx, int* n_assigned, int device);
/* UndistortKeyPoints + AssignFeaturesToGrid for a whole batch, on the device-resident output of
 * orbx_extract_batch_device (d_kps [n_frames][cap], d_counts [n_frames]) without a host round trip, on `stream`:
 * d_kps_un [n_frames][cap], d_cell_ptr [n_frames][64*48 + 1], d_cell_idx [n_frames][cap].  bounds from
 * orbf_image_bounds (they depend on the calibration only).  Two launches. */
int orbf_build_frames_device(const void* d_kps, const int32_t* d_counts, int n_frames, int cap, const float* K,
                             const float* dist, int ndist, const float* bounds, void* d_kps_un,
                             int32_t* d_cell_ptr, int32_t* d_cell_idx, void* stream);
/* vector<size_t> Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel)  R21/src/Frame.cc:332-385, for nq windows
 * at once.  Query q owns out_idx[out_ptr[q] .. out_ptr[q+1]) in the reference's order (cells ix-major, then the
 * cell's push_back order).  ORB_ERR_CAPACITY when the total exceeds cap (out_ptr is still complete). */
int orbf_features_in_area(const orb_keypoint_t* kps_un, int n, const int32_t* cell_ptr, const int32_t* cell_idx,
                          const float* bounds, const float* qx, const float* qy, const float* qr,
                          const int32_t* min_level, const int32_t* max_level, int nq, int32_t* out_ptr,
                          int32_t* out_idx, int cap, int device);

/* What Frame::isInFrustum (R21/src/Frame.cc:262-330) leaves on a MapPoint for the projection search. */
typedef struct {
    float proj_x, proj_y, proj_xr;   /* mTrackProjX, mTrackProjY, mTrackProjXR */
    float view_cos;                  /* mTrackViewCos */
    int32_t level;                   /* mnTrackScaleLevel */
    int32_t in_view;                 /* mbTrackInView && !isBad()  (:52-56) */
    int32_t obs_positive;            /* Observations() > 0, read when a later point meets the feature this one took (:82-84) */
} orbm_map_point_view_t;

/* int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
 * R21/src/ORBmatcher.cc:45-130 (Tracking::SearchLocalPoints).  Frame side: undistorted key points, descriptors,
 * mvuRight, occupied[idx] = (F.mvpMapPoints[idx] && Observations() > 0) on entry, the grid of orbf_assign_grid,
 * F.mvScaleFactors.  th_high = ORBmatcher::TH_HIGH (100).  out_feature_point[idx] = index of the map point this
 * call left in F.mvpMapPoints[idx] (-1: untouched); out_point_feature[i] = feature taken by point i or -1;
 * *n_matches = return value.  The order dependence of the reference loop is reproduced exactly. */
int orbm_search_by_projection_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right,
                                    const uint8_t* occupied, int n_f, const int32_t* cell_ptr,
                                    const int32_t* cell_idx, const float* bounds, const float* scale_factors,
                                    int n_levels, const orbm_map_point_view_t* mps, const uint8_t* desc_mp, int n_mp,
                                    float th, float nnratio, int th_high, int32_t* out_feature_point,
                                    int32_t* out_point_feature, int* n_matches, int device);

/* A projected point for the two best-only window searches below: what the reference computes before it calls
 * GetFeaturesInArea (unchanged reference statements on the caller's side). */
typedef struct {
    float u, v;            /* projection into the current frame (:1365-1366, :1505-1506) */
    float ur;              /* u - mbf * invzc (:1410); ignored by the key-frame variant */
    float angle;           /* LastFrame.mvKeysUn[i].angle / pKF->mvKeysUn[i].angle (rotation histogram) */
    int32_t octave;        /* LastFrame.mvKeys[i].octave (:1378) / pMP->PredictScale(dist3D, &CurrentFrame) (:1522) */
    int32_t valid;         /* every test before the window search passed (:1354-1373 / :1491-1520) */
    int32_t obs_positive;  /* pMP->Observations() > 0, read by later iterations (:1404-1406); ignored by the key-frame variant */
} orbm_proj_point_t;

/* int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
 * R21/src/ORBmatcher.cc:1328-1470 (Tracking::TrackWithMotionModel).  direction: 0 neither, 1 bForward, 2 bBackward
 * (:1348-1349).  occupied[f] = CurrentFrame.mvpMapPoints[f] && Observations() > 0 on entry.  out_feature_point[f]:
 * -1 untouched, -2 set to NULL by the rotation check (:1457-1461), else the index of the point left there. */
int orbm_search_by_projection_last_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right,
                                         const uint8_t* occupied, int n_f, const int32_t* cell_ptr,
                                         const int32_t* cell_idx, const float* bounds, const float* scale_factors,
                                         int n_levels, const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts,
                                         float th, int direction, int check_orientation, int th_high,
                                         int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches,
                                         int device);
/* int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound,
 *                                    const float th, const int ORBdist)
 * R21/src/ORBmatcher.cc:1472-1599 (Tracking::Relocalization).  occupied[f] = CurrentFrame.mvpMapPoints[f] != NULL. */
int orbm_search_by_projection_keyframe(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied,
                                       int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                       const float* scale_factors, int n_levels, const orbm_proj_point_t* pts,
                                       const uint8_t* desc_pts, int n_pts, float th, int orb_dist,
                                       int check_orientation, int32_t* out_feature_point, int32_t* out_point_feature,
                                       int* n_matches, int device);

/* int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints,
 *                                    vector<MapPoint*>& vpMatched, int th)
 * R21/src/ORBmatcher.cc:290-403 (LoopClosing::ComputeSim3).  pts[i]: u, v (:331-335), octave = PredictScale (:359),
 * valid = every test of :315-357 passed.  occupied[f] = vpMatched[f] != NULL on entry; th_low = TH_LOW (50).
 * grid_origin: a KeyFrame keeps its image bounds as int (R21/include/KeyFrame.h: const int mnMinX, mnMinY) and
 * KeyFrame::GetFeaturesInArea (R21/src/KeyFrame.cc:570-609) measures cells from those truncated values while the cell
 * size stays the Frame's float one; pass {(float)pKF->mnMinX, (float)pKF->mnMinY} (NULL: bounds[0], bounds[2]). */
int orbm_search_by_projection_sim3(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                   const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                   const float* scale_factors, int n_levels, const orbm_proj_point_t* pts,
                                   const uint8_t* desc_pts, int n_pts, float th, int th_low,
                                   int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches,
                                   const float* grid_origin, int device);

/* The window search inside
 *   int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)   R21/src/ORBmatcher.cc:825-975
 *       (inv_level_sigma2 = pKF->mvInvLevelSigma2: the chi-square gates of :905-931; u_right = pKF->mvuRight; pts[i].ur = :866),
 *   int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, vpPoints, float th, vpReplacePoint)        :977-1100  (inv_level_sigma2 = NULL),
 *   int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12, s12, R12, t12, th) :1102-1326 (each of its two passes).
 * Per point: the best feature of levels [l-1, l] in the window of half size th * scale_factors[l] (first wins):
 * best_idx[i] (-1: none) and best_dist[i] (256: none).  Points do not influence each other; the caller applies
 * the reference's map updates / mutual-consistency test (:945-970, :1075-1095, :1300-1323) in its original order. */
int orbm_window_best_match(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, int n_f,
                           const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                           const float* scale_factors, const float* inv_level_sigma2, int n_levels,
                           const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int32_t* best_idx,
                           int32_t* best_dist, const float* grid_origin /* as above; all three callers look up a KeyFrame */,
                           int device);

/* int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 *                                         vector<int>& vnMatches12, int windowSize)
 * R21/src/ORBmatcher.cc:405-520 (monocular initialisation).  Frame 1: undistorted key points + descriptors; frame 2:
 * the same plus its grid.  prev_xy [n1][2] = vbPrevMatched, updated in place for the matched features (:513-516);
 * out_matches12 [n1] = vnMatches12; th_low = TH_LOW (50).  The eviction / vMatchedDistance bookkeeping (:451-466) is
 * reproduced in the reference's point order. */
int orbm_search_for_initialization(const orb_keypoint_t* kps1_un, const uint8_t* desc1, int n1,
                                   const orb_keypoint_t* kps2_un, const uint8_t* desc2, int n2, const int32_t* cell_ptr,
                                   const int32_t* cell_idx, const float* bounds, float* prev_xy, int window_size,
                                   float nnratio, int check_orientation, int th_low, int32_t* out_matches12,
                                   int* n_matches, int device);

/* ---------------------------------------------------------------- BoW transform --------------- */
/* void Frame::ComputeBoW() R21/src/Frame.cc:400-407 / KeyFrame::ComputeBoW() R21/src/KeyFrame.cc:60-69:
 *     mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
 * (DBoW2::TemplatedVocabulary, R21/include/ORBVocabulary.h:31-32; DBoW2 itself is third-party and not vendored).
 * The vocabulary tree is passed as flat arrays and kept on the device: node i owns the children
 * child_idx[child_ptr[i] .. child_ptr[i+1]) (none: a leaf = a word), descriptor node_desc[i][32], and for leaves
 * word_id[i] / weight[i].  Node 0 is the root; depth_L = the vocabulary's L. */
typedef struct orbv_handle_s* orbv_handle_t;
int orbv_create(const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc, const int32_t* word_id,
                const double* weight, int n_nodes, int depth_L, int device, orbv_handle_t* out);
int orbv_destroy(orbv_handle_t h);
/* transform(feature, word, weight, &nid, levelsup) for n descriptors: greedy descent by Hamming distance, the first
 * child wins ties; out_node = the ancestor at level depth_L - levelsup (the root, 0, if that is <= 0). */
int orbv_transform(orbv_handle_t h, const uint8_t* desc, int n, int levelsup, int32_t* out_word, int32_t* out_node,
                   double* out_weight);
/* The std::map-shaped results of transform(features, v, fv, levelsup) from the per-feature triples (host side, in
 * the reference's accumulation order): BowVector = (bow_words ascending, bow_values, L1-normalised if asked),
 * FeatureVector = CSR (fv_nodes ascending, fv_ptr, fv_idx) -- the layout orbm_featvec_t views.
 * Buffers: bow_words/bow_values/fv_nodes/fv_idx [n], fv_ptr [n+1]. */
int orbv_bow_vectors(const int32_t* word, const int32_t* node, const double* weight, int n, int normalize_l1,
                     int32_t* bow_words, double* bow_values, int* n_words, int32_t* fv_nodes, int32_t* fv_ptr,
                     int32_t* fv_idx, int* n_fv_nodes);

/* ---------------------------------------------------------------- key-frame message (agent -> server) ---- */
/* The reference ships key frames as an LCM message whose key points are int16-truncated
 * (R21/include/lcmKeyFrame/lcmKeyPoint.hpp:19-31, filled R21/Examples/ROS/ORB_SLAM2/src/ros_mono.cc:2071-2077) and whose
 * descriptors travel one float per byte (lossless).  On one box the message is the device-resident output of
 * orbx_extract_batch_device exchanged with ncclAllGather; these two calls apply the wire's only lossy step so the
 * receiver sees what the reference's server sees: x, y, size, response -> (float)(int16_t)v; octave, class_id ->
 * (int16_t).  Device form: d_kps = [n_frames][cap] key points, d_counts = valid entries per frame (NULL: all cap). */
int orbw_quantize_lcm_host(orb_keypoint_t* kps, int n);
int orbw_quantize_lcm_device(void* d_kps, const int32_t* d_counts, int n_frames, int cap, void* stream);

/* The key-frame message as one contiguous device buffer for a batch of key frames (what lcmKeyFrameInfo carries for the
 * server's KeyFrame rebuild, R21/include/lcmKeyFrame/lcmKeyFrameInfo.hpp:24-145, filled ros_mono.cc:1929-2399, decoded
 * ORB_SLAM2/Examples/ROS/ORB_SLAM2/src/ros_mono.cc:230-544): per frame the key point count, mvKeys (and optionally mvKeysUn)
 * with the wire's int16 truncation applied, the 32-byte descriptors, optionally mvuRight / mvDepth and per feature the
 * map-point record {x, y, z, has} of lcmKeyFrameMapPoints (16 bytes).  Entries are compacted to the counts; the section
 * offsets depend on (n_frames, cap, flags) only.  flags = 1 (u_right + depth) | 2 (map points) | 4 (kps_un), derived from
 * the non-NULL inputs by pack and passed to unpack.  orbw_message_bytes = size of the buffer (a multiple of 16). */
size_t orbw_message_bytes(int n_frames, int cap, int flags);
int orbw_pack_keyframes_device(const void* d_kps, const void* d_kps_un, const uint8_t* d_desc, const int32_t* d_counts,
                               const float* d_u_right, const float* d_depth, const float* d_mappoints, int n_frames,
                               int cap, void* d_msg, void* stream);
/* n_msgs messages (one per sending agent, msg_stride bytes apart: the [world][slot_bytes] result of the exchange below) in ONE
 * launch; outputs are [n_msgs][n_frames][cap] (counts [n_msgs][n_frames]).  d_counts[m][f] = -1 if message m does not match
 * (n_frames, cap, flags). */
int orbw_unpack_keyframes_device(const void* d_msg, int n_msgs, size_t msg_stride, int n_frames, int cap, int flags,
                                 void* d_kps, void* d_kps_un, uint8_t* d_desc, int32_t* d_counts, float* d_u_right,
                                 float* d_depth, float* d_mappoints, void* stream);

/* ---------------------------------------------------------------- sharded map search: fused merge + exchange ---- */
/* The one exchange step of the cross-agent map search without NCCL: every rank (one process per GPU) owns a
 * symmetric record buffer, shared through CUDA IPC; orbm_knn2_exchange_device searches the rank's map shard and then ONE
 * kernel merges the rank's splits, stores its records into every rank's buffer over NVLink, publishes a flag, waits for
 * the other ranks' flags and merges the world's records: d_out [nq][4] = {d1, i1, d2, i2} of the whole map on every
 * rank, bit-identical to a single-GPU search.  Setup: orbm_peer_create on every rank (returns a 64-byte IPC handle),
 * exchange the handles with any host-side mechanism (bench.py: torch.distributed.all_gather), orbm_peer_connect with
 * the world x 64 bytes.  All ranks must make the same sequence of exchange calls.  The wait is bounded (5 s): on
 * expiry the call's output is undefined and orbm_peer_error reports 1 + the rank that did not arrive. */
typedef struct orbm_peer_s* orbm_peer_t;
int orbm_peer_create(int nq_cap, int rank, int world, int device, orbm_peer_t* out, void* ipc_handle64);
int orbm_peer_connect(orbm_peer_t p, const void* handles /* [world][64] */);
int orbm_knn2_exchange_device(orbm_peer_t p, const uint8_t* d_q, int nq, const uint8_t* d_m_shard, int64_t nm,
                              int64_t index_base, int32_t* d_out, int variant, void* stream);
/* All-gather of one message per rank (agent -> every server) over the same peer buffers, ONE kernel: stores into every
 * rank's slot over NVLink, flag, bounded wait, copy-out.  d_all = [world][slot_bytes]; the peer object must have been
 * created with nq_cap >= bytes / 16 and should not be shared with searches. */
int orbw_exchange_messages_device(orbm_peer_t p, const void* d_msg, size_t bytes, void* d_all, size_t slot_bytes,
                                  void* stream);
int orbm_peer_error(orbm_peer_t p, int* error);
int orbm_peer_destroy(orbm_peer_t p);

#ifdef __cplusplus
}
#endif
#endif /* ORBCUDA_H */
