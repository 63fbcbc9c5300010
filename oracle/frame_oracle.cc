/*
 * frame_oracle.cc -- CPU ORACLE (TEST INFRASTRUCTURE ONLY), part of liborb_oracle.so.
 *
 * Restatement of the steps right after extraction and of the projection search that consumes them
 * (SURVEY.md 8f rows 3 and 1):
 *   R21/src/Frame.cc:409-439  Frame::UndistortKeyPoints      (cv::undistortPoints, R = I, P = K)
 *   R21/src/Frame.cc:441-470  Frame::ComputeImageBounds
 *   R21/src/Frame.cc:235-250  Frame::AssignFeaturesToGrid  + :387-397 Frame::PosInGrid
 *   R21/src/Frame.cc:332-385  Frame::GetFeaturesInArea
 *   R21/src/ORBmatcher.cc:45-130  ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th)
 * cv::undistortPoints lives in OpenCV (not vendored by the reference); its published algorithm
 * (calib3d undistort: 5 fixed-point iterations in double, no tilt, P*R folded into one matrix) is restated here
 * and PINNED against cv2 4.13.0 by tests/golden/undistort_*.npz (made by tools/gen_golden_frame.py).
 * The grid / area / projection loops need Frame/MapPoint objects and cannot be compiled from the reference:
 * "parity unpinned" beyond the cv2-pinned arithmetic; they follow the cited lines statement by statement.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <algorithm>
#include <vector>

extern "C" {

/* cv::undistortPoints(src, dst, K, D, noArray(), K): K = (fx, fy, cx, cy) as float (mK is CV_32F), D = ndist floats
 * (k1, k2, p1, p2[, k3]); all arithmetic in double exactly as OpenCV orders it. */
void orc_undistort_points(const float* xy, int n, const float* K, const float* dist, int ndist, float* out) {
    double k[14] = {0};
    for (int i = 0; i < ndist && i < 14; i++) k[i] = (double)dist[i];
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // RR = P * I (a double GEMM: products with 0/1 and sums with 0 are exact)
        const double xx = fx * x + 0. * y + cx;
        const double yy = 0. * x + fy * y + cy;
        const double ww = 1. / (0. * x + 0. * y + 1.);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}

/* Frame::UndistortKeyPoints: copies everything but pt (:412-416 when k1 == 0, else :419-438). */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float* K, const float* dist, int ndist, orc_keypoint* out) {
    if (dist[0] == 0.0f) {
        for (int i = 0; i < n; i++) out[i] = kps[i];
        return;
    }
    std::vector<float> xy(2 * (size_t)n), un(2 * (size_t)n);
    for (int i = 0; i < n; i++) { xy[2 * i] = kps[i].x; xy[2 * i + 1] = kps[i].y; }
    orc_undistort_points(xy.data(), n, K, dist, ndist, un.data());
    for (int i = 0; i < n; i++) { out[i] = kps[i]; out[i].x = un[2 * i]; out[i].y = un[2 * i + 1]; }
}

/* Frame::ComputeImageBounds :441-470 -> bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY) */
void orc_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds) {
    if (dist[0] != 0.0f) {
        const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
        float u[8];
        orc_undistort_points(c, 4, K, dist, ndist, u);
        bounds[0] = std::min(u[0], u[4]);
        bounds[1] = std::max(u[2], u[6]);
        bounds[2] = std::min(u[1], u[3]);
        bounds[3] = std::max(u[5], u[7]);
    } else {
        bounds[0] = 0.f; bounds[1] = (float)cols; bounds[2] = 0.f; bounds[3] = (float)rows;
    }
}

enum { kGridCols = 64, kGridRows = 48 };   /* FRAME_GRID_COLS / FRAME_GRID_ROWS, R21/include/Frame.h:36-37 */

/* Frame::AssignFeaturesToGrid :235-250 with PosInGrid :387-397.  The grid comes back as CSR over cells in the
 * reference's mGrid[ix][iy] order (cell = ix * 48 + iy); within a cell indices ascend (push_back order). */
void orc_assign_grid(const orc_keypoint* kps_un, int n, const float* bounds, int32_t* cell_ptr, int32_t* cell_idx) {
    const float winv = (float)kGridCols / (float)(bounds[1] - bounds[0]);   /* Frame.cc:216 */
    const float hinv = (float)kGridRows / (float)(bounds[3] - bounds[2]);   /* :217 */
    std::vector<std::vector<int32_t> > grid(kGridCols * kGridRows);
    for (int i = 0; i < n; i++) {
        const int px = (int)round((kps_un[i].x - bounds[0]) * winv);
        const int py = (int)round((kps_un[i].y - bounds[2]) * hinv);
        if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
        grid[px * kGridRows + py].push_back(i);
    }
    int32_t at = 0;
    for (int c = 0; c < kGridCols * kGridRows; c++) {
        cell_ptr[c] = at;
        for (int32_t i : grid[c]) cell_idx[at++] = i;
    }
    cell_ptr[kGridCols * kGridRows] = at;
}

/* Frame::GetFeaturesInArea :332-385.  Returns the number of indices (writes at most cap). */
static int features_in_area_o(const orc_keypoint* kps_un, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                              const float* origin, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);
int orc_features_in_area(const orc_keypoint* kps_un, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                         float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    return features_in_area_o(kps_un, cell_ptr, cell_idx, bounds, nullptr, x, y, r, min_level, max_level, out, cap);
}
/* origin != NULL: KeyFrame::GetFeaturesInArea (R21/src/KeyFrame.cc:570-609) -- the key frame's mnMinX / mnMinY are ints
 * (KeyFrame.h), truncated from the Frame's float bounds, while mfGridElementWidthInv/HeightInv are the Frame's (KeyFrame.cc:33). */
static int features_in_area_o(const orc_keypoint* kps_un, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                              const float* origin, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    const float winv = (float)kGridCols / (float)(bounds[1] - bounds[0]);
    const float hinv = (float)kGridRows / (float)(bounds[3] - bounds[2]);
    const float minx = origin ? origin[0] : bounds[0], miny = origin ? origin[1] : bounds[2];
    int n = 0;
    const int nMinCellX = std::max(0, (int)floor((x - minx - r) * winv));
    if (nMinCellX >= kGridCols) return 0;
    const int nMaxCellX = std::min((int)kGridCols - 1, (int)ceil((x - minx + r) * winv));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = std::max(0, (int)floor((y - miny - r) * hinv));
    if (nMinCellY >= kGridRows) return 0;
    const int nMaxCellY = std::min((int)kGridRows - 1, (int)ceil((y - miny + r) * hinv));
    if (nMaxCellY < 0) return 0;
    const bool bCheckLevels = (min_level > 0) || (max_level >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * kGridRows + iy;
            for (int32_t j = cell_ptr[c]; j < cell_ptr[c + 1]; j++) {
                const orc_keypoint& kp = kps_un[cell_idx[j]];
                if (bCheckLevels) {
                    if (kp.octave < min_level) continue;
                    if (max_level >= 0 && kp.octave > max_level) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) {
                    if (n < cap) out[n] = cell_idx[j];
                    n++;
                }
            }
        }
    return n;
}

/* ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) :45-130.
 * Per map point (the fields Frame::isInFrustum leaves on it, Frame.cc:262-330): in_view (mbTrackInView && !isBad()),
 * proj_x/proj_y/proj_xr (mTrackProjX/Y/XR), level (mnTrackScaleLevel), view_cos (mTrackViewCos), descriptor, and
 * obs_positive (Observations() > 0, read when a LATER point meets the feature this one took, :82-84).
 * Per frame feature: undistorted key point, descriptor, u_right (mvuRight), occupied (mvpMapPoints[idx] &&
 * Observations() > 0 on entry).  out_feature_point[idx] = index of the map point left in F.mvpMapPoints[idx] by this
 * call (-1: untouched); out_point_feature[i] = feature matched by point i or -1.  Returns nmatches. */
int orc_search_by_projection_frame(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right,
                                   const uint8_t* occupied, int n_f, const int32_t* cell_ptr, const int32_t* cell_idx,
                                   const float* bounds, const float* scale_factors, const orc_map_point_view* mps,
                                   const uint8_t* desc_mp, int n_mp, float th, float nnratio, int th_high,
                                   int32_t* out_feature_point, int32_t* out_point_feature) {
    std::vector<uint8_t> blocked(occupied, occupied + n_f);
    for (int i = 0; i < n_f; i++) out_feature_point[i] = -1;
    std::vector<int32_t> cand((size_t)std::max(n_f, 1));
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int i = 0; i < n_mp; i++) {
        out_point_feature[i] = -1;
        const orc_map_point_view& mp = mps[i];
        if (!mp.in_view) continue;
        const int lvl = mp.level;
        float r = mp.view_cos > 0.998 ? 2.5 : 4.0;      /* RadiusByViewingCos :132-138 */
        if (bFactor) r *= th;
        const int nc = orc_features_in_area(kps_un, cell_ptr, cell_idx, bounds, mp.proj_x, mp.proj_y, r * scale_factors[lvl],
                                            lvl - 1, lvl, cand.data(), n_f);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            if (blocked[idx]) continue;
            if (u_right[idx] > 0) {
                const float er = fabsf(mp.proj_xr - u_right[idx]);
                if (er > r * scale_factors[lvl]) continue;
            }
            const int dist = orc_descriptor_distance(desc_mp + (size_t)i * 32, desc_f + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = kps_un[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps_un[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= th_high) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            out_feature_point[bestIdx] = i;              /* F.mvpMapPoints[bestIdx] = pMP */
            blocked[bestIdx] = mp.obs_positive ? 1 : 0;   /* what :82-84 will see for later points */
            out_point_feature[i] = bestIdx;
            nmatches++;
        }
    }
    return nmatches;
}

/* ComputeThreeMaxima R21/src/ORBmatcher.cc:1601-1642 on the histogram's bin sizes */
static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

enum { kHistoLength = 30 };   /* ORBmatcher::HISTO_LENGTH, R21/src/ORBmatcher.cc:38 */

/* Shared body of SearchByProjection(Frame&, const Frame&, th, bMono) :1328-1470 and
 * SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) :1472-1599 from the point where the projection
 * (u, v) is known.  keyframe_mode = 0: levels by direction (:1385-1390), a feature blocks when it holds a point with
 * observations (:1404-1406), stereo check (:1408-1414), threshold th_high (:1428).  keyframe_mode = 1: levels
 * [l-1, l+1] (:1531), any occupied feature blocks (:1543-1544), no stereo check, threshold ORBdist (:1557).
 * out_feature_point[f]: -1 untouched, -2 set to NULL by the rotation check, else the point left in mvpMapPoints[f]. */
static int projection_body(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied,
                           int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* origin,
                           const float* scale_factors, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts, float th,
                           int direction, int keyframe_mode, int check_orientation, int threshold,
                           int32_t* out_feature_point, int32_t* out_point_feature) {
    int nmatches = 0;
    std::vector<int> rotHist[kHistoLength];
    const float factor = 1.0f / kHistoLength;
    std::vector<uint8_t> blocked(occupied, occupied + n_f);
    for (int f = 0; f < n_f; f++) out_feature_point[f] = -1;
    std::vector<int32_t> cand((size_t)std::max(n_f, 1));
    for (int i = 0; i < n_pts; i++) {
        out_point_feature[i] = -1;
        const orc_proj_point& p = pts[i];
        if (!p.valid) continue;
        const int oct = p.octave;
        const float radius = th * scale_factors[oct];
        int nc;
        if (keyframe_mode == 2) nc = features_in_area_o(kps_un, cell_ptr, cell_idx, bounds, origin, p.u, p.v, radius, oct - 1, oct, cand.data(), n_f);
        else if (keyframe_mode) nc = orc_features_in_area(kps_un, cell_ptr, cell_idx, bounds, p.u, p.v, radius, oct - 1, oct + 1, cand.data(), n_f);
        else if (direction == 1) nc = orc_features_in_area(kps_un, cell_ptr, cell_idx, bounds, p.u, p.v, radius, oct, -1, cand.data(), n_f);
        else if (direction == 2) nc = orc_features_in_area(kps_un, cell_ptr, cell_idx, bounds, p.u, p.v, radius, 0, oct, cand.data(), n_f);
        else nc = orc_features_in_area(kps_un, cell_ptr, cell_idx, bounds, p.u, p.v, radius, oct - 1, oct + 1, cand.data(), n_f);
        if (nc == 0) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = cand[c];
            if (blocked[i2]) continue;
            if (!keyframe_mode && u_right[i2] > 0) {
                const float er = fabsf(p.ur - u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_descriptor_distance(desc_pts + (size_t)i * 32, desc_f + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= threshold) {
            out_feature_point[bestIdx2] = i;
            blocked[bestIdx2] = keyframe_mode ? 1 : (p.obs_positive ? 1 : 0);
            out_point_feature[i] = bestIdx2;
            nmatches++;
            if (check_orientation) {
                float rot = p.angle - kps_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == kHistoLength) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, kHistoLength, ind1, ind2, ind3);
        for (int i = 0; i < kHistoLength; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    out_feature_point[rotHist[i][j]] = -2;      /* mvpMapPoints[...] = NULL */
                    nmatches--;
                }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) :1328-1470.  Per last-frame
 * feature the caller supplies what :1354-1378 compute (valid = has a point, not an outlier, invzc >= 0, inside the
 * image bounds; u, v; ur = u - mbf*invzc; octave and angle of LastFrame's key point; obs_positive of its point).
 * direction: 0 neither, 1 bForward, 2 bBackward (:1348-1349). */
int orc_search_by_projection_last_frame(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right,
                                        const uint8_t* occupied, int n_f, const int32_t* cell_ptr, const int32_t* cell_idx,
                                        const float* bounds, const float* scale_factors, const orc_proj_point* pts,
                                        const uint8_t* desc_pts, int n_pts, float th, int direction, int check_orientation,
                                        int th_high, int32_t* out_feature_point, int32_t* out_point_feature) {
    return projection_body(kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds, nullptr, scale_factors, pts, desc_pts, n_pts,
                           th, direction, 0, check_orientation, th_high, out_feature_point, out_point_feature);
}

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) :1472-1599.
 * valid = the point exists, is not bad, is not in sAlreadyFound, projects inside the bounds and its distance is
 * inside the scale-invariance range (:1494-1520); octave = PredictScale (:1522); angle = pKF->mvKeysUn[i].angle;
 * occupied[f] = CurrentFrame.mvpMapPoints[f] != NULL. */
int orc_search_by_projection_keyframe(const orc_keypoint* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                      const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                      const float* scale_factors, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts,
                                      float th, int orb_dist, int check_orientation, int32_t* out_feature_point,
                                      int32_t* out_point_feature) {
    return projection_body(kps_un, desc_f, nullptr, occupied, n_f, cell_ptr, cell_idx, bounds, nullptr, scale_factors, pts, desc_pts, n_pts, th,
                           0, 1, check_orientation, orb_dist, out_feature_point, out_point_feature);
}

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th) :290-403 (loop closing) from the
 * point where (u, v) and the predicted level are known.  KeyFrame::GetFeaturesInArea (KeyFrame.cc:570-609) has no
 * level filter; the loop applies [l-1, l] itself (:364-367), which selects the same features in the same order.
 * occupied[f] = vpMatched[f] != NULL on entry; a match sets it (:381).  Threshold TH_LOW (:378), no rotation check. */
int orc_search_by_projection_sim3(const orc_keypoint* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                  const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                  const float* scale_factors, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts,
                                  float th, int th_low, int32_t* out_feature_point, int32_t* out_point_feature, const float* grid_origin) {
    return projection_body(kps_un, desc_f, nullptr, occupied, n_f, cell_ptr, cell_idx, bounds, grid_origin, scale_factors, pts, desc_pts, n_pts, th,
                           0, 2, 0, th_low, out_feature_point, out_point_feature);
}

/* The independent window search shared by Fuse(KeyFrame*, vpMapPoints, th) :825-975 (inv_level_sigma2 != NULL: the
 * chi-square gates of :905-931), Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) :977-1100 and both passes of
 * SearchBySim3 :1102-1326 (inv_level_sigma2 == NULL): best feature of levels [l-1, l] in the window, first wins.
 * No point influences another; the map updates after the search stay with the caller.  best_dist = 256 and
 * best_idx = -1 when nothing qualifies (the reference starts from 256 or INT_MAX; both fail every threshold). */
void orc_window_best_match(const orc_keypoint* kps_un, const uint8_t* desc_f, const float* u_right, int n_f,
                           const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors,
                           const float* inv_level_sigma2, const orc_proj_point* pts, const uint8_t* desc_pts, int n_pts, float th,
                           int32_t* best_idx, int32_t* best_dist, const float* grid_origin) {
    std::vector<int32_t> cand((size_t)std::max(n_f, 1));
    for (int i = 0; i < n_pts; i++) {
        best_idx[i] = -1; best_dist[i] = 256;
        const orc_proj_point& p = pts[i];
        if (!p.valid) continue;
        const int lvl = p.octave;
        const float radius = th * scale_factors[lvl];
        const float u = p.u, v = p.v, ur = p.ur;
        const int nc = features_in_area_o(kps_un, cell_ptr, cell_idx, bounds, grid_origin, u, v, radius, -1, -1, cand.data(), n_f);
        int bestDist = 256, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            const orc_keypoint& kp = kps_un[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            if (inv_level_sigma2) {
                if (u_right[idx] >= 0) {
                    const float ex = u - kp.x, ey = v - kp.y, er = ur - u_right[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = orc_descriptor_distance(desc_pts + (size_t)i * 32, desc_f + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_idx[i] = bestIdx; best_dist[i] = bestDist;
    }
}

/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) :405-520.
 * kps1_un / desc1: F1.mvKeysUn / F1.mDescriptors; frame 2 as key points + grid; prev_xy [n1][2] = vbPrevMatched (updated
 * in place for the matched features, :513-516).  out_matches12 [n1] = vnMatches12.  Returns nmatches. */
int orc_search_for_initialization(const orc_keypoint* kps1_un, const uint8_t* desc1, int n1, const orc_keypoint* kps2_un,
                                  const uint8_t* desc2, int n2, const int32_t* cell_ptr, const int32_t* cell_idx,
                                  const float* bounds, float* prev_xy, int window_size, float nnratio, int check_orientation,
                                  int th_low, int32_t* out_matches12) {
    int nmatches = 0;
    for (int i = 0; i < n1; i++) out_matches12[i] = -1;
    std::vector<int> rotHist[kHistoLength];
    const float factor = 1.0f / kHistoLength;
    std::vector<int> vMatchedDistance((size_t)n2, INT_MAX);
    std::vector<int> vnMatches21((size_t)n2, -1);
    std::vector<int32_t> cand((size_t)std::max(n2, 1));
    for (int i1 = 0; i1 < n1; i1++) {
        const orc_keypoint& kp1 = kps1_un[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        const int nc = orc_features_in_area(kps2_un, cell_ptr, cell_idx, bounds, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window_size,
                                            level1, level1, cand.data(), n2);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = cand[c];
            const int dist = orc_descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= th_low) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { out_matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                out_matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_orientation) {
                    float rot = kps1_un[i1].angle - kps2_un[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)round(rot * factor);
                    if (bin == kHistoLength) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, kHistoLength, ind1, ind2, ind3);
        for (int i = 0; i < kHistoLength; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                const int idx1 = rotHist[i][j];
                if (out_matches12[idx1] >= 0) { out_matches12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (out_matches12[i1] >= 0) { prev_xy[2 * i1] = kps2_un[out_matches12[i1]].x; prev_xy[2 * i1 + 1] = kps2_un[out_matches12[i1]].y; }
    return nmatches;
}

}  // extern "C"
