/*
 * orb_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain scalar C++ restatement of the reference front-end hot path
 *   R21 = /root/reference/ORB_SLAM2.1
 *   R21/src/ORBextractor.cc  (pyramid, cell FAST, quadtree, IC_Angle, blur, rBRIEF)
 *   R21/src/ORBmatcher.cc    (DescriptorDistance, SearchByBoW x2, SearchForTriangulation)
 *   R21/src/Frame.cc:471-645 (ComputeStereoMatches)
 * plus integer models of the OpenCV primitives those files call (resize,
 * copyMakeBorder, GaussianBlur, FAST, fastAtan2), pinned against cv2 4.13.0.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product (liborbcuda.so) never does.
 *
 * PARITY PINNING: the reference ships no tests / golden vectors.
 * This oracle is pinned two ways: (1) every primitive is checked bit-for-bit against
 * cv2 4.13.0 (tests/test_oracle_primitives.py + tests/golden/), (2) the whole
 * extractor is checked bit-for-bit against the reference's own ORBextractor.cc
 * compiled verbatim over oracle/cvshim (oracle/_ref/liborbref.so, built by
 * oracle/Makefile; tests/test_oracle_extractor.py).  The matcher loops cannot be
 * compiled from the reference (they need Frame/KeyFrame/MapPoint/DBoW2), so the
 * matcher restatement is "parity unpinned" beyond DescriptorDistance KATs.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 28-byte layout as cv::KeyPoint. */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orc_keypoint;

/* ---- OpenCV primitive models ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                          uint8_t* dst, int dw, int dh, size_t dstride);
void orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstride,
                                     uint8_t* dst, size_t dstride, int top, int bottom,
                                     int left, int right);
void orc_gaussian_blur7_sigma2(const uint8_t* src, int w, int h, size_t sstride,
                               uint8_t* dst, size_t dstride);
/* cv::FAST(img, kps, th, nms) TYPE_9_16. Returns the number found (writes at most cap). */
int orc_fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold, int nms,
                 orc_keypoint* out, int cap);
/* FAST corner score of one pixel (0 if not a corner at `threshold`). */
int orc_fast_score(const uint8_t* center, size_t stride, int threshold);
float orc_fast_atan2(float y, float x);
int orc_cv_round_f(float v);

/* ---- extractor (R21/src/ORBextractor.cc) ---- */
typedef struct orc_extractor orc_extractor;

/* trig_mode 0: a=cosf(angle), b=sinf(angle) exactly as the reference (libm, :113-114).
 * trig_mode 1: correctly rounded a=(float)cos((double)angle) (what the CUDA path uses).
 * fma_mode  0: separate fp32 mul,mul,add in the rotation (canonical, -ffp-contract=off).
 * fma_mode  1: fma(x,b, y*a) / fma(x,a, -(y*b)) -- what g++ -O3 -march=native makes of
 *              :119-120 on an FMA host. */
orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int ini_th_fast, int min_th_fast, int trig_mode,
                                    int fma_mode);
void orc_extractor_destroy(orc_extractor* e);
/* mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2, mnFeaturesPerLevel (:415-446) */
void orc_extractor_tables(const orc_extractor* e, float* sf, float* isf, float* s2, float* is2,
                          int* nfeat_per_level, int* umax16);
/* operator() (:1043-1105). Returns 0 on success; *n = number of keypoints (<= cap written). */
int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride,
                orc_keypoint* kps, uint8_t* desc, int cap, int* n);
/* stage dumps of the last orc_extract call */
int orc_level_size(const orc_extractor* e, int level, int* w, int* h);
/* with_border: copies (w+38)x(h+38) padded plane, else the w x h ROI */
int orc_get_pyramid(const orc_extractor* e, int level, int with_border, uint8_t* dst, size_t dstride);
int orc_get_blurred(const orc_extractor* e, int level, uint8_t* dst, size_t dstride);
/* FAST candidates handed to DistributeOctTree (cell-major order, coords relative to (16,16)) */
int orc_get_candidates(const orc_extractor* e, int level, int16_t* x, int16_t* y, uint8_t* score, int cap);
/* keypoints per level after octree + orientation (level coords, before the final scale) */
int orc_get_level_keypoints(const orc_extractor* e, int level, orc_keypoint* out, int cap);
/* Stand-alone DistributeOctTree (:539-763) on caller-provided candidates. */
int orc_distribute_octtree(const int16_t* x, const int16_t* y, const uint8_t* score, int n,
                           int min_x, int max_x, int min_y, int max_y, int n_features,
                           int32_t* out_index, int cap);
/* IC_Angle (:77-104) on a padded plane: center points at pixel (x,y). */
float orc_ic_angle(const uint8_t* center, size_t stride);
/* computeOrbDescriptor (:108-147) */
void orc_orb_descriptor(const uint8_t* center, size_t stride, float angle_deg, int trig_mode,
                        int fma_mode, uint8_t* desc32);

/* ---- matcher (R21/src/ORBmatcher.cc) ---- */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b); /* :1647-1663 */

/* Brute-force 2-NN with the reference update rule (:216-225): for each query the best
 * (first index on ties), its distance and the second-best distance; both start at 256. */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* m, int64_t nm, int64_t index_base,
              int32_t* best_idx, int32_t* best_d1, int32_t* best_d2, int nthreads);
/* Like orc_knn2 but also reports the index of the second best (lexicographic (d,idx) order);
 * used to check the sharded merge. */
void orc_knn2_full(const uint8_t* q, int nq, const uint8_t* m, int64_t nm, int64_t index_base,
                   int32_t* i1, int32_t* d1, int32_t* i2, int32_t* d2, int nthreads);

/* MapPoint::ComputeDistinctiveDescriptors (R21/src/MapPoint.cc:242-307) for a batch of map points: point p owns
 * the observation descriptors desc[ptr[p] .. ptr[p+1]); best[p] = index (within the point's list) of the descriptor
 * with the least median Hamming distance to the others (-1 for a point without observations). */
void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* ptr, int n_points, int32_t* best);

/* FeatureVector as CSR: node_ids ascending; node i owns idx[ptr[i]..ptr[i+1]) */
typedef struct {
    int32_t n_nodes;
    const int32_t* node_ids;
    const int32_t* ptr;
    const int32_t* idx;
} orc_featvec;

/* SearchByBoW(KeyFrame*,Frame&,...) :159-288.  kf_valid[i]!=0 <=> KF feature i has a good
 * MapPoint.  out_match_f[j] = KF feature index matched to frame feature j, or -1.
 * Returns nmatches. */
int orc_search_by_bow_kf_f(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid,
                           int n_kf, const orc_featvec* fv_kf, const uint8_t* desc_f,
                           const float* angle_f, int n_f, const orc_featvec* fv_f, float nnratio,
                           int check_ori, int32_t* out_match_f);
/* SearchByBoW(KeyFrame*,KeyFrame*,...) :522-655. out_match12[i1] = idx2 or -1. */
int orc_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                            const orc_featvec* fv1, const uint8_t* desc2, const float* angle2,
                            const uint8_t* valid2, int n2, const orc_featvec* fv2, float nnratio,
                            int check_ori, int32_t* out_match12);

typedef struct {
    float x, y;       /* mvKeysUn[i].pt */
    float angle;      /* mvKeysUn[i].angle */
    int32_t octave;   /* mvKeysUn[i].octave */
    float u_right;    /* mvuRight[i] (<0: monocular) */
    int32_t has_mp;   /* GetMapPoint(i) != NULL */
} orc_tri_feature;

/* SearchForTriangulation :657-823.  F12 row-major 3x3; (ex,ey) epipole in image 2 (:664-670);
 * scale_factors2 / level_sigma2_2 are pKF2->mvScaleFactors / mvLevelSigma2.
 * out_pairs = (idx1, idx2) sorted by idx1.  Returns nmatches. */
int orc_search_for_triangulation(const uint8_t* desc1, const orc_tri_feature* f1, int n1,
                                 const orc_featvec* fv1, const uint8_t* desc2,
                                 const orc_tri_feature* f2, int n2, const orc_featvec* fv2,
                                 const float* F12, float ex, float ey, const float* scale_factors2,
                                 const float* level_sigma2_2, int only_stereo, int check_ori,
                                 int32_t* out_pairs, int cap_pairs);

/* Frame::ComputeStereoMatches (R21/src/Frame.cc:471-645).  Pyramids are the padded planes of
 * the two extractors (ROI origin at +19,+19; plane l is (w_l+38) wide with `strides[l]`).
 * mb = mbf/fx is passed explicitly (the reference reads an uninitialised mb). */
int orc_stereo_matches(const orc_keypoint* kl, const uint8_t* dl, int nl, const orc_keypoint* kr,
                       const uint8_t* dr, int nr, int nlevels, const float* scale_factors,
                       const float* inv_scale_factors, const uint8_t* const* pyr_l,
                       const uint8_t* const* pyr_r, const int* lvl_w, const int* lvl_h,
                       const size_t* strides, float mbf, float mb, float* u_right, float* depth);

/* ---- frame side (oracle/frame_oracle.cc): the steps right after extraction + the projection search ---- */
/* cv::undistortPoints(src, dst, K, D, noArray(), K) as Frame.cc:428 calls it; K = (fx, fy, cx, cy). */
void orc_undistort_points(const float* xy, int n, const float* K, const float* dist, int ndist, float* out);
/* Frame::UndistortKeyPoints (R21/src/Frame.cc:409-439) */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float* K, const float* dist, int ndist, orc_keypoint* out);
/* Frame::ComputeImageBounds (:441-470): bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY) */
void orc_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds);
/* Frame::AssignFeaturesToGrid (:235-250): CSR over the 64 x 48 cells, cell = ix * 48 + iy; cell_ptr[3073], cell_idx[n] */
void orc_assign_grid(const orc_keypoint* kps_un, int n, const float* bounds, int32_t* cell_ptr, int32_t* cell_idx);
/* Frame::GetFeaturesInArea (:332-385) */
int orc_features_in_area(const orc_keypoint* kps_un, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                         float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);

typedef struct {
    float proj_x, proj_y, proj_xr;   /* mTrackProjX, mTrackProjY, mTrackProjXR */
    float view_cos;                  /* mTrackViewCos */
    int32_t level;                   /* mnTrackScaleLevel */
    int32_t in_view;                 /* mbTrackInView && !isBad() */
    int32_t obs_positive;            /* Observations() > 0 */
} orc_map_point_view;

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (R21/src/ORBmatcher.cc:45-130) */
int orc_search_by_projection_frame(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right,
                                   const uint8_t* occupied, int n_f, const int32_t* cell_ptr, const int32_t* cell_idx,
                                   const float* bounds, const float* scale_factors, const orc_map_point_view* mps,
                                   const uint8_t* desc_mp, int n_mp, float th, float nnratio, int th_high,
                                   int32_t* out_feature_point, int32_t* out_point_feature);

/* A projected point for the window searches below (what the reference computes before GetFeaturesInArea). */
typedef struct {
    float u, v;            /* projection into the current frame */
    float ur;              /* u - mbf * invzc (:1410) */
    float angle;           /* angle of the source key point (rotation histogram) */
    int32_t octave;        /* source key point octave (:1378) or PredictScale (:1522) */
    int32_t valid;         /* passes every test before the window search */
    int32_t obs_positive;  /* Observations() > 0 of the point (read by later iterations, :1404-1406) */
} orc_proj_point;

/* ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (R21/src/ORBmatcher.cc:1328-1470) */
int orc_search_by_projection_last_frame(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right,
                                        const uint8_t* occupied, int n_f, const int32_t* cell_ptr, const int32_t* cell_idx,
                                        const float* bounds, const float* scale_factors, const orc_proj_point* pts,
                                        const uint8_t* desc_pts, int n_pts, float th, int direction, int check_orientation,
                                        int th_high, int32_t* out_feature_point, int32_t* out_point_feature);
/* ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (R21/src/ORBmatcher.cc:1472-1599) */
int orc_search_by_projection_keyframe(const orc_keypoint* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                      const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                      const float* scale_factors, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts,
                                      float th, int orb_dist, int check_orientation, int32_t* out_feature_point,
                                      int32_t* out_point_feature);

/* ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (R21/src/ORBmatcher.cc:290-403) */
int orc_search_by_projection_sim3(const orc_keypoint* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                  const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                  const float* scale_factors, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts,
                                  float th, int th_low, int32_t* out_feature_point, int32_t* out_point_feature,
                                  const float* grid_origin /* NULL or {(float)pKF->mnMinX, (float)pKF->mnMinY}: KeyFrame.cc:575-589 */);
/* Window search of ORBmatcher::Fuse x2 (:825-975 with the chi-square gates, :977-1100) and SearchBySim3 (:1102-1326) */
void orc_window_best_match(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right, int n_f,
                           const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors,
                           const float* inv_level_sigma2, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts, float th,
                           int32_t* best_idx, int32_t* best_dist, const float* grid_origin /* as above */);

/* ORBmatcher::SearchForInitialization (R21/src/ORBmatcher.cc:405-520) */
int orc_search_for_initialization(const orc_keypoint* kps1_un, const uint8_t* desc1, int n1, const orc_keypoint* kps2_un,
                                  const uint8_t* desc2, int n2, const int32_t* cell_ptr, const int32_t* cell_idx,
                                  const float* bounds, float* prev_xy, int window_size, float nnratio, int check_orientation,
                                  int th_low, int32_t* out_matches12);

/* ---- BoW transform (oracle/bow_oracle.cc): DBoW2 TemplatedVocabulary::transform, PARITY UNPINNED (DBoW2 is not vendored) ---- */
void orc_bow_transform(const uint8_t* desc, int n, const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc,
                       const int32_t* word_id, const double* weight, int depth_L, int levelsup, int32_t* out_word,
                       int32_t* out_node, double* out_weight);
int orc_bow_vectors(const int32_t* word, const int32_t* node, const double* weight, int n, int normalize_l1, int32_t* bow_words,
                    double* bow_values, int32_t* fv_nodes, int32_t* fv_ptr, int32_t* fv_idx, int* n_nodes);

#ifdef __cplusplus
}
#endif
#endif
