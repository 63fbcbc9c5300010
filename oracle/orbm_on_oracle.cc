// orbm_on_oracle.cc -- TEST INFRASTRUCTURE ONLY.  The entry points of include/orbcuda.h that the C++ drop-in
// (cooperative-orb-slam_b200/shim/ORBmatcher.cc) calls, implemented by forwarding to the CPU restatement (orc_*,
// oracle/orb_oracle.cc + frame_oracle.cc).  Linked ONLY into oracle/_ref/libmatchshim_cpu.so, where it lets the CPU test
// suite run "reference objects -> shim ORBmatcher -> restatement" against the reference's own ORBmatcher.cc: that pins the
// restatement (and the adapters) to the reference's code without a GPU.  The product never links this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "orb_oracle.h"
#include "orbcuda.h"

// ---- optional recording (ORBM_RECORD_DIR): every call's POD inputs and outputs as .npy files, so that the scenes driven
// through the reference's objects can be replayed anywhere through the C ABI / the oracle without /root/reference
// (tools/gen_golden_matcher.py keeps them only after the scenario's results were checked against the reference's).
namespace {
int g_call = 0;
const char* rec_dir() { static const char* d = getenv("ORBM_RECORD_DIR"); return d && *d ? d : nullptr; }
void save_npy(const char* fn, const char* arg, const char* descr, const void* data, size_t rows, size_t cols, size_t elem) {
    if (!rec_dir()) return;
    char path[1024];
    snprintf(path, sizeof(path), "%s/%03d.%s.%s.npy", rec_dir(), g_call, fn, arg);
    FILE* f = fopen(path, "wb");
    if (!f) return;
    std::string shape = cols ? "(" + std::to_string(rows) + ", " + std::to_string(cols) + ")" : "(" + std::to_string(rows) + ",)";
    std::string hdr = std::string("{'descr': ") + descr + ", 'fortran_order': False, 'shape': " + shape + ", }";
    while ((10 + hdr.size() + 1) % 64) hdr += ' ';
    hdr += '\n';
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const unsigned short hl = (unsigned short)hdr.size();
    fwrite(magic, 1, 8, f); fwrite(&hl, 2, 1, f); fwrite(hdr.data(), 1, hdr.size(), f);
    if (data && rows) fwrite(data, elem, rows * (cols ? cols : 1), f);
    fclose(f);
}
#define KP_DESCR "[('x','<f4'),('y','<f4'),('size','<f4'),('angle','<f4'),('response','<f4'),('octave','<i4'),('class_id','<i4')]"
#define PROJ_DESCR "[('u','<f4'),('v','<f4'),('ur','<f4'),('angle','<f4'),('octave','<i4'),('valid','<i4'),('obs_positive','<i4')]"
#define MPV_DESCR "[('proj_x','<f4'),('proj_y','<f4'),('proj_xr','<f4'),('view_cos','<f4'),('level','<i4'),('in_view','<i4'),('obs_positive','<i4')]"
#define TRI_DESCR "[('x','<f4'),('y','<f4'),('angle','<f4'),('octave','<i4'),('u_right','<f4'),('has_mp','<i4')]"
void rec_u8(const char* fn, const char* a, const uint8_t* p, size_t n, size_t c = 0) { save_npy(fn, a, "'|u1'", p, n, c, 1); }
void rec_i32(const char* fn, const char* a, const int32_t* p, size_t n, size_t c = 0) { save_npy(fn, a, "'<i4'", p, n, c, 4); }
void rec_f32(const char* fn, const char* a, const float* p, size_t n, size_t c = 0) { save_npy(fn, a, "'<f4'", p, p ? n : 0, c, 4); }
void rec_kp(const char* fn, const char* a, const void* p, size_t n) { save_npy(fn, a, KP_DESCR, p, n, 0, 28); }
void rec_scalars(const char* fn, const float* v, size_t n) { save_npy(fn, "scalars", "'<f4'", v, n, 0, 4); }
void rec_fv(const char* fn, const char* a, const orbm_featvec_t* fv) {
    std::string s(a);
    rec_i32(fn, (s + "_ids").c_str(), fv->node_ids, fv->n_nodes);
    rec_i32(fn, (s + "_ptr").c_str(), fv->ptr, fv->n_nodes + 1);
    rec_i32(fn, (s + "_idx").c_str(), fv->idx, fv->n_nodes ? fv->ptr[fv->n_nodes] : 0);
}
void rec_grid(const char* fn, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* origin) {
    rec_i32(fn, "cell_ptr", cell_ptr, ORBF_GRID_COLS * ORBF_GRID_ROWS + 1);
    rec_i32(fn, "cell_idx", cell_idx, cell_ptr[ORBF_GRID_COLS * ORBF_GRID_ROWS]);
    rec_f32(fn, "bounds", bounds, 4);
    if (origin) rec_f32(fn, "grid_origin", origin, 2);
}
}  // namespace

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
    const char* fn = "assign_grid";
    rec_kp(fn, "kps_un", kps_un, n); rec_f32(fn, "bounds", bounds, 4);
    rec_i32(fn, "out_cell_ptr", cell_ptr, ORBF_GRID_COLS * ORBF_GRID_ROWS + 1); rec_i32(fn, "out_cell_idx", cell_idx, cell_ptr[ORBF_GRID_COLS * ORBF_GRID_ROWS]);
    g_call++;
    return ORB_OK;
}

int orbm_search_by_bow_kf_f(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid, int n_kf, const orbm_featvec_t* fv_kf,
                            const uint8_t* desc_f, const float* angle_f, int n_f, const orbm_featvec_t* fv_f, float nnratio, int check_ori,
                            int32_t* out_match_f, int* n_matches, int) {
    *n_matches = orc_search_by_bow_kf_f(desc_kf, angle_kf, kf_valid, n_kf, (const orc_featvec*)fv_kf, desc_f, angle_f, n_f,
                                        (const orc_featvec*)fv_f, nnratio, check_ori, out_match_f);
    const char* fn = "bow_kf_f";
    const float sc[3] = {nnratio, (float)check_ori, (float)*n_matches};
    rec_u8(fn, "desc_kf", desc_kf, n_kf, 32); rec_f32(fn, "angle_kf", angle_kf, n_kf); rec_u8(fn, "kf_valid", kf_valid, n_kf); rec_fv(fn, "fv_kf", fv_kf);
    rec_u8(fn, "desc_f", desc_f, n_f, 32); rec_f32(fn, "angle_f", angle_f, n_f); rec_fv(fn, "fv_f", fv_f); rec_scalars(fn, sc, 3);
    rec_i32(fn, "out_match_f", out_match_f, n_f);
    g_call++;
    return ORB_OK;
}

int orbm_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, const orbm_featvec_t* fv1,
                             const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2, const orbm_featvec_t* fv2,
                             float nnratio, int check_ori, int32_t* out_match12, int* n_matches, int) {
    *n_matches = orc_search_by_bow_kf_kf(desc1, angle1, valid1, n1, (const orc_featvec*)fv1, desc2, angle2, valid2, n2,
                                         (const orc_featvec*)fv2, nnratio, check_ori, out_match12);
    const char* fn = "bow_kf_kf";
    const float sc[3] = {nnratio, (float)check_ori, (float)*n_matches};
    rec_u8(fn, "desc1", desc1, n1, 32); rec_f32(fn, "angle1", angle1, n1); rec_u8(fn, "valid1", valid1, n1); rec_fv(fn, "fv1", fv1);
    rec_u8(fn, "desc2", desc2, n2, 32); rec_f32(fn, "angle2", angle2, n2); rec_u8(fn, "valid2", valid2, n2); rec_fv(fn, "fv2", fv2);
    rec_scalars(fn, sc, 3); rec_i32(fn, "out_match12", out_match12, n1);
    g_call++;
    return ORB_OK;
}

int orbm_search_for_triangulation(const uint8_t* desc1, const orbm_tri_feature_t* f1, int n1, const orbm_featvec_t* fv1, const uint8_t* desc2,
                                  const orbm_tri_feature_t* f2, int n2, const orbm_featvec_t* fv2, const float* F12, float ex, float ey,
                                  const float* scale_factors2, const float* level_sigma2_2, int only_stereo, int check_ori,
                                  int32_t* out_pairs, int cap_pairs, int* n_matches, int) {
    *n_matches = orc_search_for_triangulation(desc1, (const orc_tri_feature*)f1, n1, (const orc_featvec*)fv1, desc2, (const orc_tri_feature*)f2,
                                              n2, (const orc_featvec*)fv2, F12, ex, ey, scale_factors2, level_sigma2_2, only_stereo, check_ori,
                                              out_pairs, cap_pairs);
    const char* fn = "triangulation";
    const float sc[5] = {ex, ey, (float)only_stereo, (float)check_ori, (float)*n_matches};
    rec_u8(fn, "desc1", desc1, n1, 32); save_npy(fn, "f1", TRI_DESCR, f1, n1, 0, 24); rec_fv(fn, "fv1", fv1);
    rec_u8(fn, "desc2", desc2, n2, 32); save_npy(fn, "f2", TRI_DESCR, f2, n2, 0, 24); rec_fv(fn, "fv2", fv2);
    rec_f32(fn, "F12", F12, 9); rec_f32(fn, "scale_factors2", scale_factors2, 8); rec_f32(fn, "level_sigma2_2", level_sigma2_2, 8);
    rec_scalars(fn, sc, 5); rec_i32(fn, "out_pairs", out_pairs, *n_matches, 2);
    g_call++;
    return ORB_OK;
}

int orbm_search_by_projection_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied, int n_f,
                                    const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int,
                                    const orbm_map_point_view_t* mps, const uint8_t* desc_mp, int n_mp, float th, float nnratio, int th_high,
                                    int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int) {
    *n_matches = orc_search_by_projection_frame((const orc_keypoint*)kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds,
                                                scale_factors, (const orc_map_point_view*)mps, desc_mp, n_mp, th, nnratio, th_high,
                                                out_feature_point, out_point_feature);
    const char* fn = "proj_frame";
    const float sc[4] = {th, nnratio, (float)th_high, (float)*n_matches};
    rec_kp(fn, "kps_un", kps_un, n_f); rec_u8(fn, "desc_f", desc_f, n_f, 32); rec_f32(fn, "u_right", u_right, n_f); rec_u8(fn, "occupied", occupied, n_f);
    rec_grid(fn, cell_ptr, cell_idx, bounds, nullptr); rec_f32(fn, "scale_factors", scale_factors, 8);
    save_npy(fn, "mps", MPV_DESCR, mps, n_mp, 0, 28); rec_u8(fn, "desc_mp", desc_mp, n_mp, 32); rec_scalars(fn, sc, 4);
    rec_i32(fn, "out_feature_point", out_feature_point, n_f); rec_i32(fn, "out_point_feature", out_point_feature, n_mp);
    g_call++;
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
    const char* fn = "proj_last";
    const float sc[5] = {th, (float)direction, (float)check_orientation, (float)th_high, (float)*n_matches};
    rec_kp(fn, "kps_un", kps_un, n_f); rec_u8(fn, "desc_f", desc_f, n_f, 32); rec_f32(fn, "u_right", u_right, n_f); rec_u8(fn, "occupied", occupied, n_f);
    rec_grid(fn, cell_ptr, cell_idx, bounds, nullptr); rec_f32(fn, "scale_factors", scale_factors, 8);
    save_npy(fn, "pts", PROJ_DESCR, pts, n_pts, 0, 28); rec_u8(fn, "desc_pts", desc_pts, n_pts, 32); rec_scalars(fn, sc, 5);
    rec_i32(fn, "out_feature_point", out_feature_point, n_f); rec_i32(fn, "out_point_feature", out_point_feature, n_pts);
    g_call++;
    return ORB_OK;
}

int orbm_search_by_projection_keyframe(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                       const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int,
                                       const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int orb_dist,
                                       int check_orientation, int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int) {
    *n_matches = orc_search_by_projection_keyframe((const orc_keypoint*)kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors,
                                                   (const orc_proj_point*)pts, desc_pts, n_pts, th, orb_dist, check_orientation,
                                                   out_feature_point, out_point_feature);
    const char* fn = "proj_keyframe";
    const float sc[4] = {th, (float)orb_dist, (float)check_orientation, (float)*n_matches};
    rec_kp(fn, "kps_un", kps_un, n_f); rec_u8(fn, "desc_f", desc_f, n_f, 32); rec_u8(fn, "occupied", occupied, n_f);
    rec_grid(fn, cell_ptr, cell_idx, bounds, nullptr); rec_f32(fn, "scale_factors", scale_factors, 8);
    save_npy(fn, "pts", PROJ_DESCR, pts, n_pts, 0, 28); rec_u8(fn, "desc_pts", desc_pts, n_pts, 32); rec_scalars(fn, sc, 4);
    rec_i32(fn, "out_feature_point", out_feature_point, n_f); rec_i32(fn, "out_point_feature", out_point_feature, n_pts);
    g_call++;
    return ORB_OK;
}

int orbm_search_by_projection_sim3(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f, const int32_t* cell_ptr,
                                   const int32_t* cell_idx, const float* bounds, const float* scale_factors, int, const orbm_proj_point_t* pts,
                                   const uint8_t* desc_pts, int n_pts, float th, int th_low, int32_t* out_feature_point,
                                   int32_t* out_point_feature, int* n_matches, const float* grid_origin, int) {
    *n_matches = orc_search_by_projection_sim3((const orc_keypoint*)kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors,
                                               (const orc_proj_point*)pts, desc_pts, n_pts, th, th_low, out_feature_point, out_point_feature,
                                               grid_origin);
    const char* fn = "proj_sim3";
    const float sc[3] = {th, (float)th_low, (float)*n_matches};
    rec_kp(fn, "kps_un", kps_un, n_f); rec_u8(fn, "desc_f", desc_f, n_f, 32); rec_u8(fn, "occupied", occupied, n_f);
    rec_grid(fn, cell_ptr, cell_idx, bounds, grid_origin); rec_f32(fn, "scale_factors", scale_factors, 8);
    save_npy(fn, "pts", PROJ_DESCR, pts, n_pts, 0, 28); rec_u8(fn, "desc_pts", desc_pts, n_pts, 32); rec_scalars(fn, sc, 3);
    rec_i32(fn, "out_feature_point", out_feature_point, n_f); rec_i32(fn, "out_point_feature", out_point_feature, n_pts);
    g_call++;
    return ORB_OK;
}

int orbm_window_best_match(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, int n_f, const int32_t* cell_ptr,
                           const int32_t* cell_idx, const float* bounds, const float* scale_factors, const float* inv_level_sigma2, int,
                           const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int32_t* best_idx, int32_t* best_dist,
                           const float* grid_origin, int) {
    orc_window_best_match((const orc_keypoint*)kps_un, desc_f, u_right, n_f, cell_ptr, cell_idx, bounds, scale_factors, inv_level_sigma2,
                          (const orc_proj_point*)pts, desc_pts, n_pts, th, best_idx, best_dist, grid_origin);
    const char* fn = "window_best";
    const float sc[1] = {th};
    rec_kp(fn, "kps_un", kps_un, n_f); rec_u8(fn, "desc_f", desc_f, n_f, 32); if (u_right) rec_f32(fn, "u_right", u_right, n_f);
    rec_grid(fn, cell_ptr, cell_idx, bounds, grid_origin); rec_f32(fn, "scale_factors", scale_factors, 8);
    if (inv_level_sigma2) rec_f32(fn, "inv_level_sigma2", inv_level_sigma2, 8);
    save_npy(fn, "pts", PROJ_DESCR, pts, n_pts, 0, 28); rec_u8(fn, "desc_pts", desc_pts, n_pts, 32); rec_scalars(fn, sc, 1);
    rec_i32(fn, "out_best_idx", best_idx, n_pts); rec_i32(fn, "out_best_dist", best_dist, n_pts);
    g_call++;
    return ORB_OK;
}

int orbm_search_for_initialization(const orb_keypoint_t* kps1_un, const uint8_t* desc1, int n1, const orb_keypoint_t* kps2_un, const uint8_t* desc2,
                                   int n2, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, float* prev_xy, int window_size,
                                   float nnratio, int check_orientation, int th_low, int32_t* out_matches12, int* n_matches, int) {
    const char* fn = "init";
    rec_f32(fn, "prev_xy", prev_xy, n1, 2);
    *n_matches = orc_search_for_initialization((const orc_keypoint*)kps1_un, desc1, n1, (const orc_keypoint*)kps2_un, desc2, n2, cell_ptr,
                                               cell_idx, bounds, prev_xy, window_size, nnratio, check_orientation, th_low, out_matches12);
    const float sc[5] = {(float)window_size, nnratio, (float)check_orientation, (float)th_low, (float)*n_matches};
    rec_kp(fn, "kps1_un", kps1_un, n1); rec_u8(fn, "desc1", desc1, n1, 32); rec_kp(fn, "kps2_un", kps2_un, n2); rec_u8(fn, "desc2", desc2, n2, 32);
    rec_grid(fn, cell_ptr, cell_idx, bounds, nullptr); rec_scalars(fn, sc, 5);
    rec_i32(fn, "out_matches12", out_matches12, n1); rec_f32(fn, "out_prev_xy", prev_xy, n1, 2);
    g_call++;
    return ORB_OK;
}

}  // extern "C"
