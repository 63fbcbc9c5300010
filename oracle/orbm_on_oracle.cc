// orbm_on_oracle.cc -- TEST INFRASTRUCTURE ONLY.  The entry points of include/orbcuda.h that the C++ drop-in
// (cooperative-orb-slam_b200/shim/ORBmatcher.cc) calls, implemented by forwarding to the CPU restatement (orc_*,
// oracle/orb_oracle.cc + frame_oracle.cc).  Linked ONLY into oracle/_ref/libmatchshim_cpu.so, where it lets the CPU test
// suite run "reference objects -> shim ORBmatcher -> restatement" against the reference's own ORBmatcher.cc: that pins the
// restatement (and the adapters) to the reference's code without a GPU.  The product never links this file.
#include <cstring>

#include "orb_oracle.h"
#include "orbcuda.h"

static_assert(sizeof(orb_keypoint_t) == sizeof(orc_keypoint), "key point layout");
static_assert(sizeof(orbm_featvec_t) == sizeof(orc_featvec), "feature vector layout");
static_assert(sizeof(orbm_tri_feature_t) == sizeof(orc_tri_feature), "triangulation feature layout");
static_assert(sizeof(orbm_map_point_view_t) == sizeof(orc_map_point_view), "map point view layout");
static_assert(sizeof(orbm_proj_point_t) == sizeof(orc_proj_point), "projected point layout");

extern "C" {

const char* orb_last_error(void) { return "oracle backend"; }

int orb_hamming256(const void* a, const void* b) { return orc_descriptor_distance((const uint8_t*)a, (const uint8_t*)b); }

int orbf_assign_grid(const orb_keypoint_t* kps_un, int n, const float* bounds, int32_t* cell_ptr, int32_t* cell_idx, int* n_assigned, int) {
    orc_assign_grid((const orc_keypoint*)kps_un, n, bounds, cell_ptr, cell_idx);
    if (n_assigned) *n_assigned = cell_ptr[ORBF_GRID_COLS * ORBF_GRID_ROWS];
    return ORB_OK;
}

int orbm_search_by_bow_kf_f(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid, int n_kf, const orbm_featvec_t* fv_kf,
                            const uint8_t* desc_f, const float* angle_f, int n_f, const orbm_featvec_t* fv_f, float nnratio, int check_ori,
                            int32_t* out_match_f, int* n_matches, int) {
    *n_matches = orc_search_by_bow_kf_f(desc_kf, angle_kf, kf_valid, n_kf, (const orc_featvec*)fv_kf, desc_f, angle_f, n_f,
                                        (const orc_featvec*)fv_f, nnratio, check_ori, out_match_f);
    return ORB_OK;
}

int orbm_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const orbm_featvec_t* fv1,
                             const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2, const orbm_featvec_t* fv2,
                             float nnratio, int check_ori, int32_t* out_match12, int* n_matches, int) {
    *n_matches = orc_search_by_bow_kf_kf(desc1, angle1, valid1, n1, (const orc_featvec*)fv1, desc2, angle2, valid2, n2,
                                         (const orc_featvec*)fv2, nnratio, check_ori, out_match12);
    return ORB_OK;
}

int orbm_search_for_triangulation(const uint8_t* desc1, const orbm_tri_feature_t* f1, int n1, const orbm_featvec_t* fv1, const uint8_t* desc2,
                                  const orbm_tri_feature_t* f2, int n2, const orbm_featvec_t* fv2, const float* F12, float ex, float ey,
                                  const float* scale_factors2, const float* level_sigma2_2, int only_stereo, int check_ori,
                                  int32_t* out_pairs, int cap_pairs, int* n_matches, int) {
    *n_matches = orc_search_for_triangulation(desc1, (const orc_tri_feature*)f1, n1, (const orc_featvec*)fv1, desc2, (const orc_tri_feature*)f2,
                                              n2, (const orc_featvec*)fv2, F12, ex, ey, scale_factors2, level_sigma2_2, only_stereo, check_ori,
                                              out_pairs, cap_pairs);
    return ORB_OK;
}

int orbm_search_by_projection_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied, int n_f,
                                    const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int,
                                    const orbm_map_point_view_t* mps, const uint8_t* desc_mp, int n_mp, float th, float nnratio, int th_high,
                                    int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int) {
    *n_matches = orc_search_by_projection_frame((const orc_keypoint*)kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds,
                                                scale_factors, (const orc_map_point_view*)mps, desc_mp, n_mp, th, nnratio, th_high,
                                                out_feature_point, out_point_feature);
    return ORB_OK;
}

int orbm_search_by_projection_last_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied,
                                         int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                         const float* scale_factors, int, const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts,
                                         float th, int direction, int check_orientation, int th_high, int32_t* out_feature_point,
                                         int32_t* out_point_feature, int* n_matches, int) {
    *n_matches = orc_search_by_projection_last_frame((const orc_keypoint*)kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds,
                                                     scale_factors, (const orc_proj_point*)pts, desc_pts, n_pts, th, direction,
                                                     check_orientation, th_high, out_feature_point, out_point_feature);
    return ORB_OK;
}

int orbm_search_by_projection_keyframe(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                       const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int,
                                       const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int orb_dist,
                                       int check_orientation, int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int) {
    *n_matches = orc_search_by_projection_keyframe((const orc_keypoint*)kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors,
                                                   (const orc_proj_point*)pts, desc_pts, n_pts, th, orb_dist, check_orientation,
                                                   out_feature_point, out_point_feature);
    return ORB_OK;
}

int orbm_search_by_projection_sim3(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f, const int32_t* cell_ptr,
                                   const int32_t* cell_idx, const float* bounds, const float* scale_factors, int, const orbm_proj_point_t* pts,
                                   const uint8_t* desc_pts, int n_pts, float th, int th_low, int32_t* out_feature_point,
                                   int32_t* out_point_feature, int* n_matches, const float* grid_origin, int) {
    *n_matches = orc_search_by_projection_sim3((const orc_keypoint*)kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors,
                                               (const orc_proj_point*)pts, desc_pts, n_pts, th, th_low, out_feature_point, out_point_feature,
                                               grid_origin);
    return ORB_OK;
}

int orbm_window_best_match(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, int n_f, const int32_t* cell_ptr,
                           const int32_t* cell_idx, const float* bounds, const float* scale_factors, const float* inv_level_sigma2, int,
                           const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int32_t* best_idx, int32_t* best_dist,
                           const float* grid_origin, int) {
    orc_window_best_match((const orc_keypoint*)kps_un, desc_f, u_right, n_f, cell_ptr, cell_idx, bounds, scale_factors, inv_level_sigma2,
                          (const orc_proj_point*)pts, desc_pts, n_pts, th, best_idx, best_dist, grid_origin);
    return ORB_OK;
}

int orbm_search_for_initialization(const orb_keypoint_t* kps1_un, const uint8_t* desc1, int n1, const orb_keypoint_t* kps2_un, const uint8_t* desc2,
                                   int n2, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, float* prev_xy, int window_size,
                                   float nnratio, int check_orientation, int th_low, int32_t* out_matches12, int* n_matches, int) {
    *n_matches = orc_search_for_initialization((const orc_keypoint*)kps1_un, desc1, n1, (const orc_keypoint*)kps2_un, desc2, n2, cell_ptr,
                                               cell_idx, bounds, prev_xy, window_size, nnratio, check_orientation, th_low, out_matches12);
    return ORB_OK;
}

}  // extern "C"
