// frame.cu -- the steps right after extraction and the projection search that consumes them:
//   Frame::UndistortKeyPoints      R21/src/Frame.cc:409-439   (cv::undistortPoints, R = I, P = K)
//   Frame::ComputeImageBounds      R21/src/Frame.cc:441-470
//   Frame::AssignFeaturesToGrid    R21/src/Frame.cc:235-250 + PosInGrid :387-397
//   Frame::GetFeaturesInArea       R21/src/Frame.cc:332-385
//   ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th)   R21/src/ORBmatcher.cc:45-130
// All float/double arithmetic is spelled with explicit round-to-nearest intrinsics in the reference's (OpenCV's)
// evaluation order: no FMA contraction, so the results are bit-identical to the CPU code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "internal.h"
#include "orbcuda.h"

namespace orbcuda {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;   // Frame.h:36-37
constexpr int kFrameMaxFeatures = 32768;      // per-CTA shared-memory tables are sized by the feature count

struct UndistortParams {
    double fx, fy, cx, cy, ifx, ify;
    double k[12];
    int identity;      // mDistCoef[0] == 0: key points are copied (Frame.cc:411-415)
};

// ---------------------------------------------------------------- undistort: one thread per point
__device__ __forceinline__ void undistort_point(const UndistortParams& p, float fx_in, float fy_in, float& ox, float& oy) {
    const double u = (double)fx_in, v = (double)fy_in;
    double x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx);
    double y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
    const double x0 = x, y0 = y;
    const double* k = p.k;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) {
            x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx);
            y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
            break;
        }
        // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2, left to right
        const double two_xx = __dmul_rn(__dmul_rn(2.0, x), x), two_yy = __dmul_rn(__dmul_rn(2.0, y), y);
        double dx = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y), __dmul_rn(k[3], __dadd_rn(r2, two_xx)));
        dx = __dadd_rn(__dadd_rn(dx, __dmul_rn(k[8], r2)), __dmul_rn(__dmul_rn(k[9], r2), r2));
        double dy = __dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, two_yy)), __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y));
        dy = __dadd_rn(__dadd_rn(dy, __dmul_rn(k[10], r2)), __dmul_rn(__dmul_rn(k[11], r2), r2));
        x = __dmul_rn(__dsub_rn(x0, dx), icdist);
        y = __dmul_rn(__dsub_rn(y0, dy), icdist);
    }
    // RR = P * I: xx = fx*x + 0*y + cx, ww = 1/(0*x + 0*y + 1)
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(p.fx, x), __dmul_rn(0.0, y)), p.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(p.fy, y)), p.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    ox = __double2float_rn(__dmul_rn(xx, ww));
    oy = __double2float_rn(__dmul_rn(yy, ww));
}

__global__ void undistort_kernel(const orb_keypoint_t* __restrict__ in, int n, UndistortParams p, orb_keypoint_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    orb_keypoint_t kp = in[i];
    if (!p.identity) undistort_point(p, kp.x, kp.y, kp.x, kp.y);
    out[i] = kp;
}

// batch form on the extractor's device output: frame f owns kps[f*cap .. f*cap + counts[f])
__global__ void undistort_batch_kernel(const orb_keypoint_t* __restrict__ in, const int* __restrict__ counts, int cap, UndistortParams p,
                                       orb_keypoint_t* __restrict__ out) {
    const int f = blockIdx.y;
    const int n = min(counts[f], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        orb_keypoint_t kp = in[(size_t)f * cap + i];
        if (!p.identity) undistort_point(p, kp.x, kp.y, kp.x, kp.y);
        out[(size_t)f * cap + i] = kp;
    }
}

// ---------------------------------------------------------------- grid: one CTA, stable counting sort by cell
struct GridParams { float minx, miny, winv, hinv; };

__device__ __forceinline__ int pos_in_grid(const GridParams& g, float x, float y) {
    // PosInGrid: round((x - mnMinX) * mfGridElementWidthInv) in float, half away from zero
    const int px = (int)roundf(__fmul_rn(__fsub_rn(x, g.minx), g.winv));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(y, g.miny), g.hinv));
    if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) return -1;
    return px * kGridRows + py;      // mGrid[ix][iy]
}

// one CTA per frame (blockIdx.x): counts == NULL -> a single frame of n key points, else frame f has
// min(counts[f], cap) key points at kps + f*cap and writes cell_ptr + f*(kGridCells+1), cell_idx + f*cap
__global__ void __launch_bounds__(1024) grid_kernel(const orb_keypoint_t* __restrict__ kps, int n, const int* __restrict__ counts, int cap,
                                                    GridParams g, int* __restrict__ cell_ptr, int* __restrict__ cell_idx) {
    extern __shared__ int s_dyn[];
    if (counts) {
        const int f = blockIdx.x;
        n = min(counts[f], cap);
        kps += (size_t)f * cap; cell_ptr += (size_t)f * (kGridCells + 1); cell_idx += (size_t)f * cap;
    }
    int* s_count = s_dyn;                                   // [kGridCells + 1]: counts, then exclusive offsets
    short* s_cell = reinterpret_cast<short*>(s_dyn + kGridCells + 1);   // [n]
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    for (int c = tid; c <= kGridCells; c += 1024) s_count[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int c = pos_in_grid(g, kps[i].x, kps[i].y);
        s_cell[i] = (short)c;
        if (c >= 0) atomicAdd(&s_count[c], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 3 cells per thread
    const int c0 = tid * 3;
    const int a = s_count[c0], b = s_count[c0 + 1], c = s_count[c0 + 2];
    s_part[tid] = a + b + c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = tid >= off ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    const int base = s_part[tid] - (a + b + c);
    s_count[c0] = base; s_count[c0 + 1] = base + a; s_count[c0 + 2] = base + a + b;
    if (tid == 1023) s_count[kGridCells] = s_part[1023];
    __syncthreads();
    for (int k = tid; k <= kGridCells; k += 1024) cell_ptr[k] = s_count[k];
    __syncthreads();
    // stable placement (push_back order inside a cell): ONE warp walks the key points 32 at a time in index order; inside a
    // group of 32 the rank among the same-cell lanes comes from __match_any_sync, across groups from the cell's cursor in
    // s_count, which the first lane of every cell group advances.  (A rank count over all earlier key points per key point is
    // n^2 / 2 shared-memory reads with a critical path of n: 58 us for 2000 key points, 21 us this way.)
    if (tid < 32) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int ci = i < n ? (int)s_cell[i] : -2;
            const unsigned same = __match_any_sync(0xffffffffu, ci);
            const int rank = __popc(same & ((1u << tid) - 1u));
            if (ci >= 0) cell_idx[s_count[ci] + rank] = i;
            __syncwarp();
            if (ci >= 0 && rank == 0) s_count[ci] += __popc(same);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------- GetFeaturesInArea
struct AreaWindow { int x0, x1, y0, y1; bool empty; };

__device__ __forceinline__ AreaWindow area_window(const GridParams& g, float x, float y, float r) {
    AreaWindow w;
    w.empty = true;
    w.x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.minx), r), g.winv)));
    if (w.x0 >= kGridCols) return w;
    w.x1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.minx), r), g.winv)));
    if (w.x1 < 0) return w;
    w.y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.miny), r), g.hinv)));
    if (w.y0 >= kGridRows) return w;
    w.y1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.miny), r), g.hinv)));
    if (w.y1 < 0) return w;
    w.empty = false;
    return w;
}

// visits the indices GetFeaturesInArea would return, in its order; f(idx) returns false to stop early
struct KpLite { float x, y; int octave; };      // what the window walk reads of a key point (shared-memory copy)

template <class KP, class F>
__device__ __forceinline__ void for_features_in_area(const KP* __restrict__ kps, const int* __restrict__ cell_ptr,
                                                     const int* __restrict__ cell_idx, const GridParams& g, float x, float y, float r,
                                                     int min_level, int max_level, F f) {
    const AreaWindow w = area_window(g, x, y, r);
    if (w.empty) return;
    const bool check_levels = (min_level > 0) || (max_level >= 0);
    for (int ix = w.x0; ix <= w.x1; ix++)
        for (int iy = w.y0; iy <= w.y1; iy++) {
            const int c = ix * kGridRows + iy;
            const int e = cell_ptr[c + 1];
            for (int j = cell_ptr[c]; j < e; j++) {
                const int idx = cell_idx[j];
                const int octave = kps[idx].octave;
                if (check_levels) {
                    if (octave < min_level) continue;
                    if (max_level >= 0 && octave > max_level) continue;
                }
                const float distx = __fsub_rn(kps[idx].x, x), disty = __fsub_rn(kps[idx].y, y);
                if (fabsf(distx) < r && fabsf(disty) < r)
                    if (!f(idx)) return;
            }
        }
}

__global__ void area_count_kernel(const orb_keypoint_t* __restrict__ kps, const int* __restrict__ cell_ptr, const int* __restrict__ cell_idx,
                                  GridParams g, const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qr,
                                  const int* __restrict__ qmin, const int* __restrict__ qmax, int nq, int* __restrict__ counts) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int n = 0;
    for_features_in_area(kps, cell_ptr, cell_idx, g, qx[q], qy[q], qr[q], qmin[q], qmax[q], [&](int) { n++; return true; });
    counts[q] = n;
}

// exclusive scan of counts[nq] into ptr[nq + 1], one CTA
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ counts, int nq, int* __restrict__ ptr) {
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nq; base += 1024) {
        const int v = base + tid < nq ? counts[base + tid] : 0;
        s_part[tid] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int t = tid >= off ? s_part[tid - off] : 0;
            __syncthreads();
            s_part[tid] += t;
            __syncthreads();
        }
        if (base + tid < nq) ptr[base + tid] = s_carry + s_part[tid] - v;
        __syncthreads();
        if (tid == 1023) s_carry += s_part[1023];
        __syncthreads();
    }
    if (tid == 0) ptr[nq] = s_carry;
}

__global__ void area_fill_kernel(const orb_keypoint_t* __restrict__ kps, const int* __restrict__ cell_ptr, const int* __restrict__ cell_idx,
                                 GridParams g, const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qr,
                                 const int* __restrict__ qmin, const int* __restrict__ qmax, int nq, const int* __restrict__ ptr,
                                 int* __restrict__ out, int cap) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int at = ptr[q];
    for_features_in_area(kps, cell_ptr, cell_idx, g, qx[q], qy[q], qr[q], qmin[q], qmax[q], [&](int idx) {
        if (at < cap) out[at] = idx;
        at++;
        return true;
    });
}

// ---------------------------------------------------------------- SearchByProjection: window searches
// The reference walks the projected points in order; a feature taken by an earlier point is skipped by the later ones
// (ORBmatcher.cc:82-84, :1404-1406, :1543-1544), so the result depends on the order.  One CTA reproduces it in rounds.
// In a round every unresolved point i walks its window ONCE over the features that no EARLIER point has taken so far and
//   * finds its outcome as if it were its turn: best B and second best S (the two candidates the reference's best / second
//     slots would end up holding) and whether the match is accepted;
//   * claims, with atomicMin(point index), every candidate it could still TAKE later: free, within `threshold`.
// The outcome of i depends on the identity of B and S only (removing any other candidate from the walk changes neither slot),
// and an earlier unresolved point j can take a feature only if it claimed it.  So i is final as soon as neither B nor S
// carries a claim of an earlier point; the decisions are applied after a barrier (no walk reads a feature state that a
// decision of the same round has already changed).  A feature remembers WHICH point took it: a later point that became final
// in an earlier round hides the feature from the points after it only, never from an unresolved earlier one.  Two points finalised in one round never take the same feature (the
// later one would have seen the earlier one's claim on its B), the lowest unresolved point is always final, and crowded
// scenes need as many rounds as points compete for ONE feature -- the earlier rule (a point waits while ANY of its
// candidates is claimed by ANY earlier window) needed ~30 rounds and two walks per round on 4000 points over 2000 features.
struct ProjWindow {
    float x, y, r;              // window centre and half size (already multiplied by the scale factor)
    int min_level, max_level;   // GetFeaturesInArea level filter
    float ur;                   // right coordinate for the stereo check
    int flags;                  // kWinValid | kWinBlocks | kWinStereo
};
enum { kWinValid = 1, kWinBlocks = 2, kWinStereo = 4 };

__device__ __forceinline__ int hamming32(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b) {
    int d = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) d += __popc(a[w] ^ b[w]);
    return d;
}

// RATIO: best and second best with the same-level ratio rule (:99-125); otherwise best only (:1419-1424, :1551-1555)
// STAGED: key points (x, y, octave), cell_ptr and cell_idx are copied to shared memory first, so the three window
// walks per round run at shared-memory latency (frames up to kStagedMaxFeatures features; larger ones walk global memory)
constexpr int kStagedMaxFeatures = 8192;

// One point's walk of a round: its outcome over the features that are free in `blocked`, its claims into `claim`.
template <bool RATIO, class KP, class BLOCKED>
__device__ __forceinline__ void window_point_round(int i, const ProjWindow& w, const KP* __restrict__ kps, const int* __restrict__ cell_ptr,
                                                   const int* __restrict__ cell_idx, const GridParams& g, BLOCKED blocked_for_me,
                                                   const float* __restrict__ u_right, const uint32_t* __restrict__ desc_f,
                                                   const uint32_t* __restrict__ desc_p, int threshold, float nnratio, int* claim,
                                                   uint8_t* __restrict__ resolved, int2* __restrict__ tentative) {
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1, secondIdx = -1;
    const uint32_t* dp = desc_p + (size_t)i * 8;
    for_features_in_area(kps, cell_ptr, cell_idx, g, w.x, w.y, w.r, w.min_level, w.max_level, [&](int idx) {
        if (blocked_for_me(idx)) return true;
        if ((w.flags & kWinStereo) && u_right[idx] > 0) {
            const float er = fabsf(__fsub_rn(w.ur, u_right[idx]));
            if (er > w.r) return true;
        }
        const int dist = hamming32(dp, desc_f + (size_t)idx * 8);
        if (dist <= threshold) atomicMin(&claim[idx], i);       // a feature this point could take, now or later
        if (dist < bestDist) {
            bestDist2 = bestDist; bestLevel2 = bestLevel; secondIdx = bestIdx;
            bestDist = dist; bestLevel = kps[idx].octave; bestIdx = idx;
        } else if (RATIO && dist < bestDist2) {
            bestLevel2 = kps[idx].octave; bestDist2 = dist; secondIdx = idx;
        }
        return true;
    });
    if (bestDist > threshold) { resolved[i] = 1; return; }       // nothing within reach, whatever the earlier points take
    const bool accept = !(RATIO && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2));
    tentative[i] = make_int2(bestIdx | (accept ? (1 << 30) : 0), RATIO ? secondIdx : -1);
}

// The first round -- every point is unresolved, nothing is taken yet -- on the whole GPU, a WARP per point: claims into a global
// array (preset to "no claim"), outcomes into `tentative`.  The single CTA below then starts with the decisions of round 1.
// The lanes take one grid cell of the window each (a thread walking its window alone is a chain of dependent loads -- cell
// range, feature index, key point, descriptor -- per candidate: 42 us for the bench scene).  That is possible because the
// reference's best / second slots end up holding the two smallest (distance, walk position) keys of the walk -- the first
// strict minimum, and the earliest minimum among the rest -- so the candidates may be visited in any order.
constexpr int kWinCache = 16;      // candidates per point kept for the later rounds (8 bytes each)
struct Cand2 { unsigned k1, k2; int p1, p2; };      // two smallest keys (distance << 20 | position) and their (index | octave << 16)
__device__ __forceinline__ void cand2_insert(Cand2& c, unsigned k, int p) {
    if (k < c.k1) { c.k2 = c.k1; c.p2 = c.p1; c.k1 = k; c.p1 = p; }
    else if (k < c.k2) { c.k2 = k; c.p2 = p; }
}
template <bool RATIO>
__global__ void __launch_bounds__(256) window_first_round_kernel(const orb_keypoint_t* __restrict__ kps, const uint32_t* __restrict__ desc_f,
                                                                 const float* __restrict__ u_right, const uint8_t* __restrict__ occupied,
                                                                 const int* __restrict__ cell_ptr, const int* __restrict__ cell_idx, GridParams g,
                                                                 const ProjWindow* __restrict__ wins, const uint32_t* __restrict__ desc_p, int n_p,
                                                                 float nnratio, int threshold, int* __restrict__ claim_g,
                                                                 int* __restrict__ out_point_feature, uint8_t* __restrict__ resolved,
                                                                 int2* __restrict__ tentative, int2* __restrict__ cache, int* __restrict__ cache_n) {
    __shared__ int s_n[8];
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (i >= n_p) return;
    if (lane == 0) s_n[wib] = 0;
    __syncwarp();
    const ProjWindow w = wins[i];
    if (lane == 0) out_point_feature[i] = -1;
    if (!(w.flags & kWinValid)) { if (lane == 0) { resolved[i] = 1; cache_n[i] = 0; } return; }
    Cand2 best = {0xffffffffu, 0xffffffffu, -1, -1};
    const AreaWindow aw = area_window(g, w.x, w.y, w.r);
    if (!aw.empty) {
        const bool check_levels = (w.min_level > 0) || (w.max_level >= 0);
        const uint32_t* dp = desc_p + (size_t)i * 8;
        const int ny = aw.y1 - aw.y0 + 1, ncells = (aw.x1 - aw.x0 + 1) * ny;
        int base = 0;                                   // walk position of the first entry of this group of 32 cells
        for (int c0 = 0; c0 < ncells; c0 += 32) {
            const int cl = c0 + lane;
            int beg = 0, cnt = 0;
            if (cl < ncells) {
                const int c = (aw.x0 + cl / ny) * kGridRows + aw.y0 + cl % ny;      // the walk's order: ix outer, iy inner
                beg = cell_ptr[c]; cnt = cell_ptr[c + 1] - beg;
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int pos0 = base + incl - cnt;
            base += __shfl_sync(0xffffffffu, incl, 31);
            for (int j = 0; j < cnt; j++) {
                const int idx = cell_idx[beg + j];
                const int octave = kps[idx].octave;
                if (check_levels) {
                    if (octave < w.min_level) continue;
                    if (w.max_level >= 0 && octave > w.max_level) continue;
                }
                const float distx = __fsub_rn(kps[idx].x, w.x), disty = __fsub_rn(kps[idx].y, w.y);
                if (!(fabsf(distx) < w.r && fabsf(disty) < w.r)) continue;
                if (occupied[idx]) continue;
                if ((w.flags & kWinStereo) && u_right[idx] > 0) {
                    const float er = fabsf(__fsub_rn(w.ur, u_right[idx]));
                    if (er > w.r) continue;
                }
                const int dist = hamming32(dp, desc_f + (size_t)idx * 8);
                if (dist <= threshold) atomicMin(&claim_g[idx], i);
                const unsigned key = ((unsigned)dist << 20) | (unsigned)(pos0 + j);
                const int payload = idx | (octave << 16);
                cand2_insert(best, key, payload);
                // the later rounds (window_search_kernel) re-evaluate this point from the cached candidates: the static tests
                // (window, level, occupied, stereo) and the distance are done; only "who has taken it since" changes
                const int slot = atomicAdd(&s_n[wib], 1);
                if (slot < kWinCache) cache[(size_t)i * kWinCache + slot] = make_int2((int)key, payload);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned k1 = __shfl_xor_sync(0xffffffffu, best.k1, o), k2 = __shfl_xor_sync(0xffffffffu, best.k2, o);
        const int p1 = __shfl_xor_sync(0xffffffffu, best.p1, o), p2 = __shfl_xor_sync(0xffffffffu, best.p2, o);
        cand2_insert(best, k1, p1);
        cand2_insert(best, k2, p2);
    }
    __syncwarp();
    if (lane) return;
    cache_n[i] = s_n[wib];          // more than kWinCache: the later rounds walk the window again
    const int bestDist = best.p1 >= 0 ? (int)(best.k1 >> 20) : 256, bestDist2 = best.p2 >= 0 ? (int)(best.k2 >> 20) : 256;
    if (bestDist > threshold) { resolved[i] = 1; return; }
    resolved[i] = 0;
    const int bestLevel = best.p1 >> 16, bestLevel2 = best.p2 >= 0 ? best.p2 >> 16 : -1;
    const bool accept = !(RATIO && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2));
    tentative[i] = make_int2((best.p1 & 0xffff) | (accept ? (1 << 30) : 0), RATIO && best.p2 >= 0 ? (best.p2 & 0xffff) : -1);
}

template <bool RATIO, bool STAGED>
__global__ void __launch_bounds__(1024) window_search_kernel(const orb_keypoint_t* __restrict__ kps_g, const uint32_t* __restrict__ desc_f,
                                                             const float* __restrict__ u_right, const uint8_t* __restrict__ occupied, int n_f,
                                                             const int* __restrict__ cell_ptr_g, const int* __restrict__ cell_idx_g, GridParams g,
                                                             const ProjWindow* __restrict__ wins, const uint32_t* __restrict__ desc_p, int n_p,
                                                             float nnratio, int threshold, int* __restrict__ out_feature_point,
                                                             int* __restrict__ out_point_feature, uint8_t* __restrict__ resolved,
                                                             int2* __restrict__ tentative, const int* __restrict__ first_round_claims,
                                                             const int2* __restrict__ cache, const int* __restrict__ cache_n,
                                                             int* __restrict__ taker_g, int* __restrict__ out_nmatches) {
    extern __shared__ int s_dyn[];
    int* s_claim = s_dyn;                                             // [n_f]
    int* s_cell_idx = s_claim + n_f;                                  // [n_f]            (STAGED)
    int* s_cell_ptr = s_cell_idx + (STAGED ? n_f : 0);                // [kGridCells + 1] (STAGED)
    KpLite* s_kp = reinterpret_cast<KpLite*>(s_cell_ptr + (STAGED ? kGridCells + 1 : 0));      // [n_f] (STAGED)
    int* s_taker = STAGED ? reinterpret_cast<int*>(s_kp + n_f) : taker_g;                      // [n_f] see below (global memory for large frames)
    __shared__ int s_left, s_matches;
    const int tid = threadIdx.x;
    for (int f = tid; f < n_f; f += 1024) {
        // the point that took the feature and blocks it: -1 = occupied from the start (blocked for every point), "none" otherwise.
        // Point i skips a feature only if its taker is EARLIER than i: a later point that was final in an earlier round must not
        // change what i sees at its turn.
        s_taker[f] = occupied[f] ? -1 : 0x7fffffff; out_feature_point[f] = -1;
        if (STAGED) {
            s_cell_idx[f] = cell_idx_g[f];
            KpLite k; k.x = kps_g[f].x; k.y = kps_g[f].y; k.octave = kps_g[f].octave;
            s_kp[f] = k;
        }
    }
    if (STAGED)
        for (int c = tid; c <= kGridCells; c += 1024) s_cell_ptr[c] = cell_ptr_g[c];
    if (!first_round_claims)
        for (int i = tid; i < n_p; i += 1024) {
            out_point_feature[i] = -1;
            resolved[i] = (wins[i].flags & kWinValid) ? 0 : 1;
        }
    if (tid == 0) s_matches = 0;
    const int* cell_ptr = STAGED ? s_cell_ptr : cell_ptr_g;
    const int* cell_idx = STAGED ? s_cell_idx : cell_idx_g;
    for (int round = 0;; round++) {
        __syncthreads();
        if (round == 0 && first_round_claims) {
            // round 1 was walked by window_first_round_kernel
            for (int f = tid; f < n_f; f += 1024) s_claim[f] = first_round_claims[f];
            if (tid == 0) s_left = 0;
        } else {
            for (int f = tid; f < n_f; f += 1024) s_claim[f] = 0x7fffffff;
            if (tid == 0) s_left = 0;
            __syncthreads();
            for (int i = tid; i < n_p; i += 1024) {
                if (resolved[i]) continue;
                const int nc = first_round_claims ? cache_n[i] : kWinCache + 1;
                if (nc <= kWinCache) {
                    // from the candidates the first round cached: no window walk, no descriptor loads
                    Cand2 best = {0xffffffffu, 0xffffffffu, -1, -1};
                    const int2* c = cache + (size_t)i * kWinCache;
                    for (int k = 0; k < nc; k++) {
                        const int2 e = c[k];
                        const int idx = e.y & 0xffff;
                        if (s_taker[idx] < i) continue;
                        if ((int)((unsigned)e.x >> 20) <= threshold) atomicMin(&s_claim[idx], i);
                        cand2_insert(best, (unsigned)e.x, e.y);
                    }
                    const int bestDist = best.p1 >= 0 ? (int)(best.k1 >> 20) : 256, bestDist2 = best.p2 >= 0 ? (int)(best.k2 >> 20) : 256;
                    if (bestDist > threshold) { resolved[i] = 1; continue; }
                    const int bestLevel = best.p1 >> 16, bestLevel2 = best.p2 >= 0 ? best.p2 >> 16 : -1;
                    const bool accept = !(RATIO && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2));
                    tentative[i] = make_int2((best.p1 & 0xffff) | (accept ? (1 << 30) : 0), RATIO && best.p2 >= 0 ? (best.p2 & 0xffff) : -1);
                    continue;
                }
                auto blocked_for_me = [&](int idx) { return s_taker[idx] < i; };
                if (STAGED) window_point_round<RATIO>(i, wins[i], s_kp, cell_ptr, cell_idx, g, blocked_for_me, u_right, desc_f, desc_p, threshold, nnratio, s_claim, resolved, tentative);
                else window_point_round<RATIO>(i, wins[i], kps_g, cell_ptr, cell_idx, g, blocked_for_me, u_right, desc_f, desc_p, threshold, nnratio, s_claim, resolved, tentative);
            }
        }
        __syncthreads();
        for (int i = tid; i < n_p; i += 1024) {
            if (resolved[i]) continue;
            const int2 t = tentative[i];
            const int bestIdx = t.x & ~(1 << 30);
            if (s_claim[bestIdx] < i || (t.y >= 0 && s_claim[t.y] < i)) { atomicAdd(&s_left, 1); continue; }
            resolved[i] = 1;
            if (t.x & (1 << 30)) {
                out_feature_point[bestIdx] = i;
                if (wins[i].flags & kWinBlocks) s_taker[bestIdx] = i;
                out_point_feature[i] = bestIdx;
                atomicAdd(&s_matches, 1);
            }
        }
        __syncthreads();
        if (s_left == 0) break;
    }
    if (tid == 0) *out_nmatches = s_matches;
}

// ---------------------------------------------------------------- Fuse x2 / SearchBySim3: independent window search
// One thread per projected point: the best feature of levels [l-1, l] inside the window, first wins; with
// inv_level_sigma2 the chi-square gates of ORBmatcher.cc:905-931 (7.8 with a right coordinate, 5.99 without; the
// float product is compared with the double literal as the reference does).
__global__ void window_best_kernel(const orb_keypoint_t* __restrict__ kps, const uint32_t* __restrict__ desc_f, const float* __restrict__ u_right,
                                   const int* __restrict__ cell_ptr, const int* __restrict__ cell_idx, GridParams g,
                                   const float* __restrict__ scale_factors, const float* __restrict__ inv_level_sigma2,
                                   const orbm_proj_point_t* __restrict__ pts, const uint32_t* __restrict__ desc_p, int n_p, float th,
                                   int* __restrict__ best_idx, int* __restrict__ best_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_p) return;
    const orbm_proj_point_t p = pts[i];
    int bestDist = 256, bestIdx = -1;
    if (p.valid) {
        const int lvl = p.octave;
        const float radius = __fmul_rn(th, scale_factors[lvl]);
        const uint32_t* dp = desc_p + (size_t)i * 8;
        for_features_in_area(kps, cell_ptr, cell_idx, g, p.u, p.v, radius, -1, -1, [&](int idx) {
            const int kl = kps[idx].octave;
            if (kl < lvl - 1 || kl > lvl) return true;
            if (inv_level_sigma2) {
                const float ex = __fsub_rn(p.u, kps[idx].x), ey = __fsub_rn(p.v, kps[idx].y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                const float kr = u_right[idx];
                if (kr >= 0) {
                    const float er = __fsub_rn(p.ur, kr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, inv_level_sigma2[kl]) > 7.8) return true;
                } else if ((double)__fmul_rn(e2, inv_level_sigma2[kl]) > 5.99) return true;
            }
            const int dist = hamming32(dp, desc_f + (size_t)idx * 8);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            return true;
        });
    }
    best_idx[i] = bestIdx; best_dist[i] = bestDist;
}

// ---------------------------------------------------------------- SearchForInitialization (ORBmatcher.cc:405-520)
// Same order problem, different state: a frame-2 feature remembers the distance of its current match
// (vMatchedDistance) and only yields to a strictly better one; a new match evicts the previous owner (:459-463).
// Rounds exactly as in window_search_kernel; the per-feature state lives in global memory (matched_dist, owner).
__global__ void __launch_bounds__(1024) init_search_kernel(const orb_keypoint_t* __restrict__ kps1, const uint32_t* __restrict__ desc1, int n1,
                                                           const orb_keypoint_t* __restrict__ kps2, const uint32_t* __restrict__ desc2, int n2,
                                                           const int* __restrict__ cell_ptr, const int* __restrict__ cell_idx, GridParams g,
                                                           const float* __restrict__ prev_xy, float window, float nnratio, int th_low,
                                                           int* __restrict__ matched_dist, int* __restrict__ owner, int* __restrict__ matches12,
                                                           int* __restrict__ ever_matched, uint8_t* __restrict__ resolved,
                                                           int* __restrict__ out_nmatches) {
    extern __shared__ int s_dyn[];
    int* s_claim = s_dyn;      // [n2]
    __shared__ int s_left, s_matches;
    const int tid = threadIdx.x;
    for (int f = tid; f < n2; f += 1024) { matched_dist[f] = 0x7fffffff; owner[f] = -1; }
    for (int i = tid; i < n1; i += 1024) {
        matches12[i] = -1; ever_matched[i] = -1;
        resolved[i] = kps1[i].octave > 0 ? 1 : 0;      // :421-423
    }
    if (tid == 0) s_matches = 0;
    while (true) {
        __syncthreads();
        for (int f = tid; f < n2; f += 1024) s_claim[f] = 0x7fffffff;
        if (tid == 0) s_left = 0;
        __syncthreads();
        for (int i = tid; i < n1; i += 1024) {
            if (resolved[i]) continue;
            for_features_in_area(kps2, cell_ptr, cell_idx, g, prev_xy[2 * i], prev_xy[2 * i + 1], window, 0, 0,
                                 [&](int idx) { atomicMin(&s_claim[idx], i); return true; });
        }
        __syncthreads();
        for (int i = tid; i < n1; i += 1024) {
            if (resolved[i]) continue;
            bool safe = true;
            int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
            const uint32_t* d1 = desc1 + (size_t)i * 8;
            for_features_in_area(kps2, cell_ptr, cell_idx, g, prev_xy[2 * i], prev_xy[2 * i + 1], window, 0, 0, [&](int idx) {
                if (s_claim[idx] < i) { safe = false; return false; }
                const int dist = hamming32(d1, desc2 + (size_t)idx * 8);
                if (matched_dist[idx] <= dist) return true;
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = idx; }
                else if (dist < bestDist2) bestDist2 = dist;
                return true;
            });
            if (!safe) { atomicAdd(&s_left, 1); continue; }
            resolved[i] = 1;
            if (bestDist <= th_low && (float)bestDist < __fmul_rn((float)bestDist2, nnratio)) {
                const int prev = owner[bestIdx2];
                if (prev >= 0) { matches12[prev] = -1; atomicSub(&s_matches, 1); }
                matches12[i] = bestIdx2;
                owner[bestIdx2] = i;
                matched_dist[bestIdx2] = bestDist;
                ever_matched[i] = bestIdx2;      // the rotation histogram keeps evicted points too (:469-479)
                atomicAdd(&s_matches, 1);
            }
        }
        __syncthreads();
        if (s_left == 0) break;
    }
    if (tid == 0) *out_nmatches = s_matches;
}

static GridParams make_grid_params(const float* bounds, const float* origin = nullptr);

// uploads the frame and the windows, runs the search, downloads feature -> point, point -> feature and the count
static int run_window_search(bool ratio, const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied,
                             int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* origin,
                             const std::vector<ProjWindow>& wins,
                             const uint8_t* desc_p, float nnratio, int threshold, int32_t* out_feature_point, int32_t* out_point_feature,
                             int* n_matches, int device) {
    const int n_p = (int)wins.size();
    MatchCtx& cx = match_ctx();
    const size_t kb = (size_t)n_f * sizeof(orb_keypoint_t), pb = (size_t)(kGridCells + 1) * 4;
    const size_t need = kb + (size_t)n_f * (32 + 4 + 1 + 4 + 4 + 4 + 4) + pb + (size_t)n_p * (sizeof(ProjWindow) + 32 + 4 + 1 + 8 + kWinCache * 8 + 4) + 4 + 32 * 256;
    if (!cx.begin(device, need, need)) return ORB_ERR_CUDA;
    std::vector<float> no_stereo;
    if (!u_right) { no_stereo.assign(n_f, -1.0f); u_right = no_stereo.data(); }
    const orb_keypoint_t* d_k = (const orb_keypoint_t*)cx.upload(kps_un, kb);
    const uint32_t* d_df = (const uint32_t*)cx.upload(desc_f, (size_t)n_f * 32);
    const float* d_ur = (const float*)cx.upload(u_right, (size_t)n_f * 4);
    const uint8_t* d_occ = (const uint8_t*)cx.upload(occupied, (size_t)n_f);
    const int* d_cp = (const int*)cx.upload(cell_ptr, pb);
    const int* d_ci = (const int*)cx.upload(cell_idx, (size_t)n_f * 4);
    const ProjWindow* d_w = (const ProjWindow*)cx.upload(wins.data(), (size_t)n_p * sizeof(ProjWindow));
    const uint32_t* d_dp = (const uint32_t*)cx.upload(desc_p, (size_t)n_p * 32);
    int* d_fp = (int*)cx.dalloc((size_t)n_f * 4); int* d_pf = (int*)cx.dalloc((size_t)n_p * 4);
    int* d_nm = (int*)cx.dalloc(4);                 // right behind d_fp and d_pf: the three downloads leave as one copy
    uint8_t* d_res = (uint8_t*)cx.dalloc((size_t)n_p);
    int2* d_tent = (int2*)cx.dalloc((size_t)n_p * sizeof(int2));
    int* d_claim = (int*)cx.dalloc((size_t)n_f * 4);
    int* d_taker = (int*)cx.dalloc((size_t)n_f * 4);
    int2* d_cache = (int2*)cx.dalloc((size_t)n_p * kWinCache * sizeof(int2));
    int* d_cache_n = (int*)cx.dalloc((size_t)n_p * 4);
    if (!d_cache || !d_cache_n || !d_taker || !d_k || !d_df || !d_ur || !d_occ || !d_cp || !d_ci || !d_w || !d_dp || !d_fp || !d_pf || !d_res || !d_nm || !d_tent || !d_claim) return ORB_ERR_CUDA;
    const bool staged = n_f <= kStagedMaxFeatures;
    static DeviceOnce once_configured;
    if (!once_configured.run([&] {
            const int big = kFrameMaxFeatures * 4 + 16, small = kStagedMaxFeatures * (4 + 4 + 12 + 4) + (kGridCells + 1) * 4 + 16;
            return cuda_ok(cudaFuncSetAttribute(window_search_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big), "cudaFuncSetAttribute") &&
                   cuda_ok(cudaFuncSetAttribute(window_search_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big), "cudaFuncSetAttribute") &&
                   cuda_ok(cudaFuncSetAttribute(window_search_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small), "cudaFuncSetAttribute") &&
                   cuda_ok(cudaFuncSetAttribute(window_search_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small), "cudaFuncSetAttribute");
        })) return ORB_ERR_CUDA;
    const size_t smem = staged ? (size_t)n_f * (4 + 4 + 12 + 4) + (kGridCells + 1) * 4 + 16 : (size_t)n_f * 4 + 16;
    const GridParams g = make_grid_params(bounds, origin);
    // round 1 on the whole GPU when there are enough points to fill more than one SM (ORBCUDA_WS_FIRST=0: everything in the one CTA)
    static const bool first_on_grid = [] { const char* e = getenv("ORBCUDA_WS_FIRST"); return e ? atoi(e) != 0 : true; }();
    const int* d_first = nullptr;
    if (first_on_grid && n_p >= 1024 && n_f > 0) {
        ORB_CUDA_TRY(cudaMemsetAsync(d_claim, 0x7f, (size_t)n_f * 4, cx.s()));      // 0x7f7f7f7f: above every point index
        if (ratio) window_first_round_kernel<true><<<(n_p + 7) / 8, 256, 0, cx.s()>>>(d_k, d_df, d_ur, d_occ, d_cp, d_ci, g, d_w, d_dp, n_p, nnratio, threshold, d_claim, d_pf, d_res, d_tent, d_cache, d_cache_n);
        else window_first_round_kernel<false><<<(n_p + 7) / 8, 256, 0, cx.s()>>>(d_k, d_df, d_ur, d_occ, d_cp, d_ci, g, d_w, d_dp, n_p, nnratio, threshold, d_claim, d_pf, d_res, d_tent, d_cache, d_cache_n);
        d_first = d_claim;
    }
#define ORB_LAUNCH_WS(R, S) window_search_kernel<R, S><<<1, 1024, smem, cx.s()>>>(d_k, d_df, d_ur, d_occ, n_f, d_cp, d_ci, g, d_w, d_dp, n_p, nnratio, threshold, d_fp, d_pf, d_res, d_tent, d_first, d_cache, d_cache_n, d_taker, d_nm)
    if (ratio) { if (staged) ORB_LAUNCH_WS(true, true); else ORB_LAUNCH_WS(true, false); }
    else { if (staged) ORB_LAUNCH_WS(false, true); else ORB_LAUNCH_WS(false, false); }
#undef ORB_LAUNCH_WS
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(out_feature_point, d_fp, (size_t)n_f * 4) || !cx.download(out_point_feature, d_pf, (size_t)n_p * 4) ||
        !cx.download(n_matches, d_nm, 4) || !cx.finish()) return ORB_ERR_CUDA;
    return ORB_OK;
}

// rotation consistency of the best-only searches (:1431-1466, :1561-1596): matches outside the three dominant bins of
// the angle-difference histogram are set to NULL; every histogram entry removed decrements the count
static void rotation_check(const orb_keypoint_t* kps_un, const orbm_proj_point_t* pts, int n_p, const int32_t* point_feature,
                           int32_t* feature_point, int* n_matches) {
    constexpr int kHisto = 30;      // ORBmatcher::HISTO_LENGTH
    const float factor = 1.0f / kHisto;
    int cnt[kHisto] = {0};
    std::vector<int> bin(n_p, -1);
    for (int i = 0; i < n_p; i++) {
        if (point_feature[i] < 0) continue;
        float rot = pts[i].angle - kps_un[point_feature[i]].angle;
        if (rot < 0.0) rot += 360.0f;
        int b = (int)roundf(rot * factor);
        if (b == kHisto) b = 0;
        bin[i] = b; cnt[b]++;
    }
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;      // ComputeThreeMaxima :1601-1642
    for (int i = 0; i < kHisto; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int i = 0; i < n_p; i++) {
        if (bin[i] < 0 || bin[i] == ind1 || bin[i] == ind2 || bin[i] == ind3) continue;
        feature_point[point_feature[i]] = -2;
        (*n_matches)--;
    }
}

static bool make_undistort_params(const float* K, const float* dist, int ndist, UndistortParams* p) {
    if (!K || !dist || ndist < 4 || ndist > 12) return false;
    p->fx = K[0]; p->fy = K[1]; p->cx = K[2]; p->cy = K[3];
    p->ifx = 1. / p->fx; p->ify = 1. / p->fy;
    for (int i = 0; i < 12; i++) p->k[i] = i < ndist ? (double)dist[i] : 0.0;
    p->identity = dist[0] == 0.0f;
    return true;
}

// origin: a KeyFrame looks its grid up from integer-truncated bounds (R21/include/KeyFrame.h: const int mnMinX, mnMinY;
// KeyFrame.cc:575-589) while the cell size is still the Frame's float one (mfGridElementWidthInv is copied, KeyFrame.cc:33)
static GridParams make_grid_params(const float* bounds, const float* origin) {
    GridParams g;
    g.minx = origin ? origin[0] : bounds[0]; g.miny = origin ? origin[1] : bounds[2];
    g.winv = (float)kGridCols / (float)(bounds[1] - bounds[0]);      // Frame.cc:216
    g.hinv = (float)kGridRows / (float)(bounds[3] - bounds[2]);      // :217
    return g;
}

}  // namespace orbcuda

using namespace orbcuda;

extern "C" {

int orbf_undistort_keypoints(const orb_keypoint_t* kps, int n, const float* K, const float* dist, int ndist, orb_keypoint_t* out, int device) {
    UndistortParams p;
    if (n < 0 || (n && (!kps || !out)) || !make_undistort_params(K, dist, ndist, &p)) { set_error("orbf_undistort_keypoints: bad arguments (K = fx,fy,cx,cy; 4..12 distortion coefficients)"); return ORB_ERR_ARG; }
    if (n == 0) return ORB_OK;
    MatchCtx& cx = match_ctx();
    const size_t bytes = (size_t)n * sizeof(orb_keypoint_t);
    if (!cx.begin(device, 2 * bytes + 1024, 2 * bytes + 1024)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_in = (const orb_keypoint_t*)cx.upload(kps, bytes);
    orb_keypoint_t* d_out = (orb_keypoint_t*)cx.dalloc(bytes);
    if (!d_in || !d_out) return ORB_ERR_CUDA;
    undistort_kernel<<<(n + 255) / 256, 256, 0, cx.s()>>>(d_in, n, p, d_out);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(out, d_out, bytes) || !cx.finish()) return ORB_ERR_CUDA;
    return ORB_OK;
}

int orbf_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds, int device) {
    if (!bounds || cols <= 0 || rows <= 0 || !dist) { set_error("orbf_image_bounds: bad arguments"); return ORB_ERR_ARG; }
    if (dist[0] == 0.0f) {      // Frame.cc:463-468
        bounds[0] = 0.f; bounds[1] = (float)cols; bounds[2] = 0.f; bounds[3] = (float)rows;
        return ORB_OK;
    }
    orb_keypoint_t c[4] = {}, u[4];
    c[1].x = (float)cols; c[2].y = (float)rows; c[3].x = (float)cols; c[3].y = (float)rows;
    const int rc = orbf_undistort_keypoints(c, 4, K, dist, ndist, u, device);
    if (rc) return rc;
    bounds[0] = std::min(u[0].x, u[2].x);      // :456-459
    bounds[1] = std::max(u[1].x, u[3].x);
    bounds[2] = std::min(u[0].y, u[1].y);
    bounds[3] = std::max(u[2].y, u[3].y);
    return ORB_OK;
}

int orbf_assign_grid(const orb_keypoint_t* kps_un, int n, const float* bounds, int32_t* cell_ptr, int32_t* cell_idx, int* n_assigned, int device) {
    if (n < 0 || n > kFrameMaxFeatures || !bounds || !cell_ptr || (n && (!kps_un || !cell_idx))) { set_error("orbf_assign_grid: bad arguments (at most %d key points)", kFrameMaxFeatures); return ORB_ERR_ARG; }
    MatchCtx& cx = match_ctx();
    const size_t kb = (size_t)std::max(n, 1) * sizeof(orb_keypoint_t), pb = (size_t)(kGridCells + 1) * 4, ib = (size_t)std::max(n, 1) * 4;
    if (!cx.begin(device, kb + pb + ib + 2048, kb + pb + ib + 2048)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_k = (const orb_keypoint_t*)cx.upload(kps_un, (size_t)n * sizeof(orb_keypoint_t));
    int* d_ptr = (int*)cx.dalloc(pb);
    int* d_idx = (int*)cx.dalloc(ib);
    if ((n && !d_k) || !d_ptr || !d_idx) return ORB_ERR_CUDA;
    const size_t smem = (size_t)(kGridCells + 1) * 4 + (size_t)n * 2 + 16;
    static DeviceOnce once_configured;
    if (!once_configured.run([&] {
            return cuda_ok(cudaFuncSetAttribute(grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kGridCells + 1) * 4 + kFrameMaxFeatures * 2 + 16), "cudaFuncSetAttribute");
        })) return ORB_ERR_CUDA;
    grid_kernel<<<1, 1024, smem, cx.s()>>>(d_k, n, nullptr, 0, make_grid_params(bounds), d_ptr, d_idx);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(cell_ptr, d_ptr, pb)) return ORB_ERR_CUDA;
    if (n && !cx.download(cell_idx, d_idx, (size_t)n * 4)) return ORB_ERR_CUDA;
    if (!cx.finish()) return ORB_ERR_CUDA;
    if (n_assigned) *n_assigned = cell_ptr[kGridCells];
    return ORB_OK;
}

int orbf_build_frame(const orb_keypoint_t* kps, int n, const float* K, const float* dist, int ndist, const float* bounds,
                     orb_keypoint_t* out_kps_un, int32_t* cell_ptr, int32_t* cell_idx, int* n_assigned, int device) {
    UndistortParams p;
    if (n < 0 || n > kFrameMaxFeatures || !bounds || !cell_ptr || (n && (!kps || !out_kps_un || !cell_idx)) || !make_undistort_params(K, dist, ndist, &p)) {
        set_error("orbf_build_frame: bad arguments (at most %d key points; K = fx,fy,cx,cy; 4..12 distortion coefficients)", kFrameMaxFeatures);
        return ORB_ERR_ARG;
    }
    MatchCtx& cx = match_ctx();
    const size_t kb = (size_t)std::max(n, 1) * sizeof(orb_keypoint_t), pb = (size_t)(kGridCells + 1) * 4, ib = (size_t)std::max(n, 1) * 4;
    if (!cx.begin(device, 2 * kb + pb + ib + 2048, 2 * kb + pb + ib + 2048)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_in = (const orb_keypoint_t*)cx.upload(kps, (size_t)n * sizeof(orb_keypoint_t));
    orb_keypoint_t* d_un = (orb_keypoint_t*)cx.dalloc(kb);      // the three results lie side by side: one download
    int* d_ptr = (int*)cx.dalloc(pb);
    int* d_idx = (int*)cx.dalloc(ib);
    if ((n && !d_in) || !d_un || !d_ptr || !d_idx) return ORB_ERR_CUDA;
    const size_t smem = (size_t)(kGridCells + 1) * 4 + (size_t)n * 2 + 16;
    static DeviceOnce once_configured;
    if (!once_configured.run([&] {
            return cuda_ok(cudaFuncSetAttribute(grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kGridCells + 1) * 4 + kFrameMaxFeatures * 2 + 16), "cudaFuncSetAttribute");
        })) return ORB_ERR_CUDA;
    if (n) undistort_kernel<<<(n + 255) / 256, 256, 0, cx.s()>>>(d_in, n, p, d_un);
    grid_kernel<<<1, 1024, smem, cx.s()>>>(d_un, n, nullptr, 0, make_grid_params(bounds), d_ptr, d_idx);
    ORB_CUDA_TRY(cudaGetLastError());
    if (n && !cx.download(out_kps_un, d_un, (size_t)n * sizeof(orb_keypoint_t))) return ORB_ERR_CUDA;
    if (!cx.download(cell_ptr, d_ptr, pb)) return ORB_ERR_CUDA;
    if (n && !cx.download(cell_idx, d_idx, (size_t)n * 4)) return ORB_ERR_CUDA;
    if (!cx.finish()) return ORB_ERR_CUDA;
    if (n_assigned) *n_assigned = cell_ptr[kGridCells];
    return ORB_OK;
}

int orbf_build_frames_device(const void* d_kps, const int32_t* d_counts, int n_frames, int cap, const float* K, const float* dist, int ndist,
                             const float* bounds, void* d_kps_un, int32_t* d_cell_ptr, int32_t* d_cell_idx, void* stream) {
    UndistortParams p;
    if (!d_kps || !d_counts || n_frames < 0 || cap <= 0 || cap > kFrameMaxFeatures || !bounds || !d_kps_un || !d_cell_ptr || !d_cell_idx ||
        !make_undistort_params(K, dist, ndist, &p)) {
        set_error("orbf_build_frames_device: bad arguments (cap <= %d)", kFrameMaxFeatures);
        return ORB_ERR_ARG;
    }
    if (n_frames == 0) return ORB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    static DeviceOnce once_configured;
    if (!once_configured.run([&] {
            return cuda_ok(cudaFuncSetAttribute(grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kGridCells + 1) * 4 + kFrameMaxFeatures * 2 + 16), "cudaFuncSetAttribute");
        })) return ORB_ERR_CUDA;
    undistort_batch_kernel<<<dim3((cap + 255) / 256, n_frames), 256, 0, s>>>((const orb_keypoint_t*)d_kps, d_counts, cap, p, (orb_keypoint_t*)d_kps_un);
    grid_kernel<<<n_frames, 1024, (size_t)(kGridCells + 1) * 4 + (size_t)cap * 2 + 16, s>>>((const orb_keypoint_t*)d_kps_un, 0, d_counts, cap,
                                                                                         make_grid_params(bounds), d_cell_ptr, d_cell_idx);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbf_features_in_area(const orb_keypoint_t* kps_un, int n, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                          const float* qx, const float* qy, const float* qr, const int32_t* min_level, const int32_t* max_level, int nq,
                          int32_t* out_ptr, int32_t* out_idx, int cap, int device) {
    if (n < 0 || nq < 0 || cap < 0 || !bounds || !cell_ptr || !out_ptr || (cap && !out_idx) || (nq && (!qx || !qy || !qr || !min_level || !max_level)) ||
        (n && (!kps_un || !cell_idx))) { set_error("orbf_features_in_area: bad arguments"); return ORB_ERR_ARG; }
    if (nq == 0) { out_ptr[0] = 0; return ORB_OK; }
    MatchCtx& cx = match_ctx();
    const size_t kb = (size_t)std::max(n, 1) * sizeof(orb_keypoint_t), pb = (size_t)(kGridCells + 1) * 4, ib = (size_t)std::max(n, 1) * 4;
    const size_t qb = (size_t)nq * 4, ob = (size_t)std::max(cap, 1) * 4;
    const size_t need = kb + pb + ib + 7 * qb + 4 + ob + 16 * 256;
    if (!cx.begin(device, need, need)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_k = (const orb_keypoint_t*)cx.upload(kps_un, (size_t)n * sizeof(orb_keypoint_t));
    const int* d_cp = (const int*)cx.upload(cell_ptr, pb);
    const int* d_ci = (const int*)cx.upload(cell_idx, (size_t)n * 4);
    const float* d_qx = (const float*)cx.upload(qx, qb); const float* d_qy = (const float*)cx.upload(qy, qb);
    const float* d_qr = (const float*)cx.upload(qr, qb);
    const int* d_mn = (const int*)cx.upload(min_level, qb); const int* d_mx = (const int*)cx.upload(max_level, qb);
    int* d_cnt = (int*)cx.dalloc(qb); int* d_ptr = (int*)cx.dalloc(qb + 4); int* d_out = (int*)cx.dalloc(ob);
    if ((n && (!d_k || !d_ci)) || !d_cp || !d_qx || !d_qy || !d_qr || !d_mn || !d_mx || !d_cnt || !d_ptr || !d_out) return ORB_ERR_CUDA;
    const GridParams g = make_grid_params(bounds);
    area_count_kernel<<<(nq + 127) / 128, 128, 0, cx.s()>>>(d_k, d_cp, d_ci, g, d_qx, d_qy, d_qr, d_mn, d_mx, nq, d_cnt);
    scan_kernel<<<1, 1024, 0, cx.s()>>>(d_cnt, nq, d_ptr);
    area_fill_kernel<<<(nq + 127) / 128, 128, 0, cx.s()>>>(d_k, d_cp, d_ci, g, d_qx, d_qy, d_qr, d_mn, d_mx, nq, d_ptr, d_out, cap);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(out_ptr, d_ptr, qb + 4)) return ORB_ERR_CUDA;
    if (cap && !cx.download(out_idx, d_out, ob)) return ORB_ERR_CUDA;
    if (!cx.finish()) return ORB_ERR_CUDA;
    if (out_ptr[nq] > cap) { set_error("orbf_features_in_area: %d indices, capacity %d (out_ptr is complete, out_idx truncated)", out_ptr[nq], cap); return ORB_ERR_CAPACITY; }
    return ORB_OK;
}

static bool check_frame_args(const char* who, const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                             const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int n_levels,
                             const void* pts, const uint8_t* desc_p, int n_p, int32_t* out_feature_point, int32_t* out_point_feature,
                             int* n_matches) {
    if (n_f < 0 || n_f > kFrameMaxFeatures || n_p < 0 || n_levels <= 0 || !bounds || !cell_ptr || !scale_factors || !n_matches ||
        (n_f && (!kps_un || !desc_f || !occupied || !cell_idx || !out_feature_point)) || (n_p && (!pts || !desc_p || !out_point_feature))) {
        set_error("%s: bad arguments (at most %d frame features)", who, kFrameMaxFeatures);
        return false;
    }
    return true;
}

int orbm_search_by_projection_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied, int n_f,
                                    const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors, int n_levels,
                                    const orbm_map_point_view_t* mps, const uint8_t* desc_mp, int n_mp, float th, float nnratio, int th_high,
                                    int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int device) {
    if (!check_frame_args("orbm_search_by_projection_frame", kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors, n_levels, mps,
                          desc_mp, n_mp, out_feature_point, out_point_feature, n_matches) || (n_f && !u_right)) {
        if (n_f && !u_right) set_error("orbm_search_by_projection_frame: u_right is required");
        return ORB_ERR_ARG;
    }
    std::vector<ProjWindow> wins(n_mp);
    const bool factor = th != 1.0f;
    for (int i = 0; i < n_mp; i++) {
        const orbm_map_point_view_t& mp = mps[i];
        ProjWindow& w = wins[i];
        w.flags = 0;
        if (!mp.in_view) continue;
        if (mp.level < 0 || mp.level >= n_levels) { set_error("orbm_search_by_projection_frame: map point %d predicts level %d of %d", i, mp.level, n_levels); return ORB_ERR_ARG; }
        float r = (double)mp.view_cos > 0.998 ? 2.5f : 4.0f;      // RadiusByViewingCos :132-138
        if (factor) r *= th;                                        // :64-65
        w.x = mp.proj_x; w.y = mp.proj_y; w.r = r * scale_factors[mp.level];      // :68
        w.min_level = mp.level - 1; w.max_level = mp.level; w.ur = mp.proj_xr;
        w.flags = kWinValid | kWinStereo | (mp.obs_positive ? kWinBlocks : 0);
    }
    *n_matches = 0;
    for (int f = 0; f < n_f; f++) out_feature_point[f] = -1;
    for (int i = 0; i < n_mp; i++) out_point_feature[i] = -1;
    if (n_f == 0 || n_mp == 0) return ORB_OK;
    return run_window_search(true, kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds, nullptr, wins, desc_mp, nnratio, th_high,
                             out_feature_point, out_point_feature, n_matches, device);
}

static int best_only_search(const char* who, bool keyframe_mode, int level_up, const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right,
                            const uint8_t* occupied, int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* origin,
                            const float* scale_factors, int n_levels, const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th,
                            int direction, int check_orientation, int threshold, int32_t* out_feature_point, int32_t* out_point_feature,
                            int* n_matches, int device) {
    if (!check_frame_args(who, kps_un, desc_f, occupied, n_f, cell_ptr, cell_idx, bounds, scale_factors, n_levels, pts, desc_pts, n_pts,
                          out_feature_point, out_point_feature, n_matches)) return ORB_ERR_ARG;
    if (direction < 0 || direction > 2 || (!keyframe_mode && n_f && !u_right)) { set_error("%s: bad direction / u_right", who); return ORB_ERR_ARG; }
    std::vector<ProjWindow> wins(n_pts);
    for (int i = 0; i < n_pts; i++) {
        const orbm_proj_point_t& p = pts[i];
        ProjWindow& w = wins[i];
        w.flags = 0;
        if (!p.valid) continue;
        const int oct = p.octave;
        if (oct < 0 || oct >= n_levels) { set_error("%s: point %d has octave %d of %d", who, i, oct, n_levels); return ORB_ERR_ARG; }
        w.x = p.u; w.y = p.v; w.r = th * scale_factors[oct];      // :1381, :1529
        if (keyframe_mode) { w.min_level = oct - 1; w.max_level = oct + level_up; }                  // :1531 (+1), :364-367 (+0)
        else if (direction == 0) { w.min_level = oct - 1; w.max_level = oct + 1; }                   // :1390
        else if (direction == 1) { w.min_level = oct; w.max_level = -1; }                            // :1386
        else { w.min_level = 0; w.max_level = oct; }                                                 // :1388
        w.ur = p.ur;
        w.flags = kWinValid | (keyframe_mode ? kWinBlocks : (kWinStereo | (p.obs_positive ? kWinBlocks : 0)));
    }
    *n_matches = 0;
    for (int f = 0; f < n_f; f++) out_feature_point[f] = -1;
    for (int i = 0; i < n_pts; i++) out_point_feature[i] = -1;
    if (n_f == 0 || n_pts == 0) return ORB_OK;
    const int rc = run_window_search(false, kps_un, desc_f, keyframe_mode ? nullptr : u_right, occupied, n_f, cell_ptr, cell_idx, bounds, origin, wins,
                                     desc_pts, 0.f, threshold, out_feature_point, out_point_feature, n_matches, device);
    if (rc) return rc;
    if (check_orientation) rotation_check(kps_un, pts, n_pts, out_point_feature, out_feature_point, n_matches);
    return ORB_OK;
}

int orbm_search_by_projection_last_frame(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, const uint8_t* occupied,
                                         int n_f, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds,
                                         const float* scale_factors, int n_levels, const orbm_proj_point_t* pts, const uint8_t* desc_pts,
                                         int n_pts, float th, int direction, int check_orientation, int th_high,
                                         int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, int device) {
    return best_only_search("orbm_search_by_projection_last_frame", false, 1, kps_un, desc_f, u_right, occupied, n_f, cell_ptr, cell_idx, bounds, nullptr,
                            scale_factors, n_levels, pts, desc_pts, n_pts, th, direction, check_orientation, th_high, out_feature_point,
                            out_point_feature, n_matches, device);
}

int orbm_search_by_projection_keyframe(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f,
                                       const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, const float* scale_factors,
                                       int n_levels, const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int orb_dist,
                                       int check_orientation, int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches,
                                       int device) {
    return best_only_search("orbm_search_by_projection_keyframe", true, 1, kps_un, desc_f, nullptr, occupied, n_f, cell_ptr, cell_idx, bounds, nullptr,
                            scale_factors, n_levels, pts, desc_pts, n_pts, th, 0, check_orientation, orb_dist, out_feature_point,
                            out_point_feature, n_matches, device);
}

int orbm_search_by_projection_sim3(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const uint8_t* occupied, int n_f, const int32_t* cell_ptr,
                                   const int32_t* cell_idx, const float* bounds, const float* scale_factors, int n_levels,
                                   const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int th_low,
                                   int32_t* out_feature_point, int32_t* out_point_feature, int* n_matches, const float* grid_origin, int device) {
    return best_only_search("orbm_search_by_projection_sim3", true, 0, kps_un, desc_f, nullptr, occupied, n_f, cell_ptr, cell_idx, bounds, grid_origin,
                            scale_factors, n_levels, pts, desc_pts, n_pts, th, 0, 0, th_low, out_feature_point, out_point_feature, n_matches,
                            device);
}

int orbm_window_best_match(const orb_keypoint_t* kps_un, const uint8_t* desc_f, const float* u_right, int n_f, const int32_t* cell_ptr,
                           const int32_t* cell_idx, const float* bounds, const float* scale_factors, const float* inv_level_sigma2, int n_levels,
                           const orbm_proj_point_t* pts, const uint8_t* desc_pts, int n_pts, float th, int32_t* best_idx, int32_t* best_dist,
                           const float* grid_origin, int device) {
    if (n_f < 0 || n_pts < 0 || n_levels <= 0 || !bounds || !cell_ptr || !scale_factors || (n_f && (!kps_un || !desc_f || !cell_idx)) ||
        (n_pts && (!pts || !desc_pts || !best_idx || !best_dist)) || (inv_level_sigma2 && n_f && !u_right)) {
        set_error("orbm_window_best_match: bad arguments (u_right is required with inv_level_sigma2)");
        return ORB_ERR_ARG;
    }
    for (int i = 0; i < n_pts; i++) {
        if (pts[i].valid && (pts[i].octave < 0 || pts[i].octave >= n_levels)) { set_error("orbm_window_best_match: point %d predicts level %d of %d", i, pts[i].octave, n_levels); return ORB_ERR_ARG; }
        best_idx[i] = -1; best_dist[i] = 256;
    }
    if (n_f == 0 || n_pts == 0) return ORB_OK;
    MatchCtx& cx = match_ctx();
    const size_t kb = (size_t)n_f * sizeof(orb_keypoint_t), pb = (size_t)(kGridCells + 1) * 4;
    const size_t need = kb + (size_t)n_f * (32 + 4 + 4) + pb + (size_t)n_levels * 8 + (size_t)n_pts * (sizeof(orbm_proj_point_t) + 32 + 8) + 20 * 256;
    if (!cx.begin(device, need, need)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_k = (const orb_keypoint_t*)cx.upload(kps_un, kb);
    const uint32_t* d_df = (const uint32_t*)cx.upload(desc_f, (size_t)n_f * 32);
    const float* d_ur = inv_level_sigma2 ? (const float*)cx.upload(u_right, (size_t)n_f * 4) : nullptr;
    const int* d_cp = (const int*)cx.upload(cell_ptr, pb);
    const int* d_ci = (const int*)cx.upload(cell_idx, (size_t)n_f * 4);
    const float* d_sf = (const float*)cx.upload(scale_factors, (size_t)n_levels * 4);
    const float* d_is = inv_level_sigma2 ? (const float*)cx.upload(inv_level_sigma2, (size_t)n_levels * 4) : nullptr;
    const orbm_proj_point_t* d_p = (const orbm_proj_point_t*)cx.upload(pts, (size_t)n_pts * sizeof(orbm_proj_point_t));
    const uint32_t* d_dp = (const uint32_t*)cx.upload(desc_pts, (size_t)n_pts * 32);
    int* d_bi = (int*)cx.dalloc((size_t)n_pts * 4); int* d_bd = (int*)cx.dalloc((size_t)n_pts * 4);
    if (!d_k || !d_df || !d_cp || !d_ci || !d_sf || !d_p || !d_dp || !d_bi || !d_bd || (inv_level_sigma2 && (!d_ur || !d_is))) return ORB_ERR_CUDA;
    window_best_kernel<<<(n_pts + 127) / 128, 128, 0, cx.s()>>>(d_k, d_df, d_ur, d_cp, d_ci, make_grid_params(bounds, grid_origin), d_sf, d_is, d_p, d_dp, n_pts, th,
                                                                  d_bi, d_bd);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(best_idx, d_bi, (size_t)n_pts * 4) || !cx.download(best_dist, d_bd, (size_t)n_pts * 4) || !cx.finish()) return ORB_ERR_CUDA;
    return ORB_OK;
}

int orbm_search_for_initialization(const orb_keypoint_t* kps1_un, const uint8_t* desc1, int n1, const orb_keypoint_t* kps2_un, const uint8_t* desc2,
                                   int n2, const int32_t* cell_ptr, const int32_t* cell_idx, const float* bounds, float* prev_xy, int window_size,
                                   float nnratio, int check_orientation, int th_low, int32_t* out_matches12, int* n_matches, int device) {
    if (n1 < 0 || n2 < 0 || n2 > kFrameMaxFeatures || !bounds || !cell_ptr || !n_matches || (n1 && (!kps1_un || !desc1 || !prev_xy || !out_matches12)) ||
        (n2 && (!kps2_un || !desc2 || !cell_idx))) {
        set_error("orbm_search_for_initialization: bad arguments (at most %d features in frame 2)", kFrameMaxFeatures);
        return ORB_ERR_ARG;
    }
    *n_matches = 0;
    for (int i = 0; i < n1; i++) out_matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORB_OK;
    MatchCtx& cx = match_ctx();
    const size_t pb = (size_t)(kGridCells + 1) * 4;
    const size_t need = (size_t)n1 * (sizeof(orb_keypoint_t) + 32 + 8 + 4 + 4 + 1) + (size_t)n2 * (sizeof(orb_keypoint_t) + 32 + 4 + 4 + 4) + pb + 4 + 24 * 256;
    if (!cx.begin(device, need, need)) return ORB_ERR_CUDA;
    const orb_keypoint_t* d_k1 = (const orb_keypoint_t*)cx.upload(kps1_un, (size_t)n1 * sizeof(orb_keypoint_t));
    const uint32_t* d_d1 = (const uint32_t*)cx.upload(desc1, (size_t)n1 * 32);
    const orb_keypoint_t* d_k2 = (const orb_keypoint_t*)cx.upload(kps2_un, (size_t)n2 * sizeof(orb_keypoint_t));
    const uint32_t* d_d2 = (const uint32_t*)cx.upload(desc2, (size_t)n2 * 32);
    const int* d_cp = (const int*)cx.upload(cell_ptr, pb);
    const int* d_ci = (const int*)cx.upload(cell_idx, (size_t)n2 * 4);
    const float* d_xy = (const float*)cx.upload(prev_xy, (size_t)n1 * 8);
    int* d_md = (int*)cx.dalloc((size_t)n2 * 4); int* d_ow = (int*)cx.dalloc((size_t)n2 * 4);
    int* d_m12 = (int*)cx.dalloc((size_t)n1 * 4); int* d_em = (int*)cx.dalloc((size_t)n1 * 4);
    uint8_t* d_res = (uint8_t*)cx.dalloc((size_t)n1); int* d_nm = (int*)cx.dalloc(4);
    if (!d_k1 || !d_d1 || !d_k2 || !d_d2 || !d_cp || !d_ci || !d_xy || !d_md || !d_ow || !d_m12 || !d_em || !d_res || !d_nm) return ORB_ERR_CUDA;
    static DeviceOnce once_configured;
    if (!once_configured.run([&] {
            return cuda_ok(cudaFuncSetAttribute(init_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrameMaxFeatures * 4 + 16), "cudaFuncSetAttribute");
        })) return ORB_ERR_CUDA;
    init_search_kernel<<<1, 1024, (size_t)n2 * 4 + 16, cx.s()>>>(d_k1, d_d1, n1, d_k2, d_d2, n2, d_cp, d_ci, make_grid_params(bounds), d_xy,
                                                                    (float)window_size, nnratio, th_low, d_md, d_ow, d_m12, d_em, d_res, d_nm);
    ORB_CUDA_TRY(cudaGetLastError());
    std::vector<int32_t> ever(n1);
    if (!cx.download(out_matches12, d_m12, (size_t)n1 * 4) || !cx.download(ever.data(), d_em, (size_t)n1 * 4) || !cx.download(n_matches, d_nm, 4) ||
        !cx.finish()) return ORB_ERR_CUDA;
    if (check_orientation) {      // :484-506 -- the histogram holds every point that ever matched, in point order
        constexpr int kHisto = 30;
        const float factor = 1.0f / kHisto;
        int cnt[kHisto] = {0};
        std::vector<int> bin(n1, -1);
        for (int i = 0; i < n1; i++) {
            if (ever[i] < 0) continue;
            float rot = kps1_un[i].angle - kps2_un[ever[i]].angle;
            if (rot < 0.0) rot += 360.0f;
            int b = (int)roundf(rot * factor);
            if (b == kHisto) b = 0;
            bin[i] = b; cnt[b]++;
        }
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < kHisto; i++) {
            const int c = cnt[i];
            if (c > max1) { max3 = max2; max2 = max1; max1 = c; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (c > max2) { max3 = max2; max2 = c; ind3 = ind2; ind2 = i; }
            else if (c > max3) { max3 = c; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < n1; i++) {
            if (bin[i] < 0 || bin[i] == ind1 || bin[i] == ind2 || bin[i] == ind3) continue;
            if (out_matches12[i] >= 0) { out_matches12[i] = -1; (*n_matches)--; }
        }
    }
    for (int i = 0; i < n1; i++)      // :513-516
        if (out_matches12[i] >= 0) { prev_xy[2 * i] = kps2_un[out_matches12[i]].x; prev_xy[2 * i + 1] = kps2_un[out_matches12[i]].y; }
    return ORB_OK;
}

}  // extern "C"
