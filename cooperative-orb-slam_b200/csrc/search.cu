// search.cu -- K8/K9: the reference's candidate loops on the GPU.
//   orbm_search_by_bow_kf_f        ORBmatcher::SearchByBoW(KeyFrame*,Frame&,...)      R21/src/ORBmatcher.cc:159-288
//   orbm_search_by_bow_kf_kf       ORBmatcher::SearchByBoW(KeyFrame*,KeyFrame*,...)   R21/src/ORBmatcher.cc:522-655
//   orbm_search_for_triangulation  ORBmatcher::SearchForTriangulation               R21/src/ORBmatcher.cc:657-823
//   orbm_stereo_matches            Frame::ComputeStereoMatches                      R21/src/Frame.cc:471-645
// The vocabulary-node walk, rotation histogram (ComputeThreeMaxima :1601-1642) and the stereo median cut
// are tiny and stay on the host; all Hamming / SAD work runs in the kernels below.  SearchByBoW's greedy
// "skip features already matched" is node-local (a feature index lives in exactly one FeatureVector
// node), so nodes run in parallel -- one warp each -- and only the walk inside a node is sequential.
#include "internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

namespace orbcuda {

constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // R21/src/ORBmatcher.cc:37-39

struct NodePair { int a0, a1, b0, b1; };

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4* __restrict__ b) {
    const uint4 b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
__device__ __forceinline__ unsigned warp_min(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One warp per shared vocabulary node.  mode 0: KF -> Frame (accept d1 <= TH_LOW, exclusion = matchB >= 0);
// mode 1: KF -> KF (accept d1 < TH_LOW, both sides need valid map points, exclusion = matched flag).
__global__ void __launch_bounds__(128) bow_kernel(const uint4* __restrict__ descA, const uint8_t* __restrict__ validA,
                                                  const int* __restrict__ idxA, const uint4* __restrict__ descB,
                                                  const uint8_t* __restrict__ validB, const int* __restrict__ idxB,
                                                  const NodePair* __restrict__ pairs, int n_pairs, float ratio, int mode,
                                                  volatile int* matchB, int* __restrict__ matchA) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const NodePair np = pairs[w];
    for (int p = np.a0; p < np.a1; p++) {
        const int ia = idxA[p];
        if (!validA[ia]) continue;
        const uint4 a0 = descA[2 * (size_t)ia], a1 = descA[2 * (size_t)ia + 1];
        int d1 = 256, d2 = 256, pos1 = 0xfffff;
        for (int r = np.b0 + lane; r < np.b1; r += 32) {
            const int ib = idxB[r];
            if (matchB[ib] >= 0) continue;
            if (mode == 1 && !validB[ib]) continue;
            const int d = hamming256(a0, a1, descB + 2 * (size_t)ib);
            if (d < d1) { d2 = d1; d1 = d; pos1 = r - np.b0; }
            else if (d < d2) { d2 = d; }
        }
        // warp merge: best = lexicographic min (distance, list position); second counts duplicates
        const unsigned key = ((unsigned)d1 << 20) | (unsigned)pos1;
        const unsigned best = warp_min(key);
        const int second = (int)warp_min((unsigned)(key == best ? d2 : d1));
        const int bd = (int)(best >> 20);
        const bool accept = (mode == 0 ? bd <= TH_LOW : bd < TH_LOW) && ((float)bd < __fmul_rn(ratio, (float)second));
        if (accept && lane == 0) {
            const int ib = idxB[np.b0 + (int)(best & 0xfffff)];
            matchB[ib] = ia;
            if (matchA) matchA[ia] = ib;
        }
        __syncwarp();
    }
}

struct TriFeat { float x, y, angle; int octave; float u_right; int has_mp; };
struct TriItem { int ia, b0, b1; };

// One warp per unmatched key point of KF1: scan its vocabulary node in KF2 (rows are independent because the
// reference never sets vbMatched2, :677,:725).  Winner = qualifying candidate of minimal distance, last on ties.
__global__ void __launch_bounds__(128) triangulation_kernel(const uint4* __restrict__ desc1, const TriFeat* __restrict__ f1,
                                                            const uint4* __restrict__ desc2, const TriFeat* __restrict__ f2,
                                                            const int* __restrict__ idx2, const TriItem* __restrict__ items,
                                                            int n_items, const float* __restrict__ F, float ex, float ey,
                                                            const float* __restrict__ sf2, const float* __restrict__ sig2,
                                                            int only_stereo, int* __restrict__ match12) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_items) return;
    const TriItem it = items[w];
    const TriFeat k1 = f1[it.ia];
    const bool stereo1 = k1.u_right >= 0;
    const uint4 a0 = desc1[2 * (size_t)it.ia], a1 = desc1[2 * (size_t)it.ia + 1];
    // epipolar line l = x1' F12 (CheckDistEpipolarLine :140-157)
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[0]), __fmul_rn(k1.y, F[3])), F[6]);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[1]), __fmul_rn(k1.y, F[4])), F[7]);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[2]), __fmul_rn(k1.y, F[5])), F[8]);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    unsigned key = 0xffffffffu;
    for (int r = it.b0 + lane; r < it.b1; r += 32) {
        const int ib = idx2[r];
        const TriFeat k2 = f2[ib];
        if (k2.has_mp) continue;
        const bool stereo2 = k2.u_right >= 0;
        if (only_stereo && !stereo2) continue;
        const int d = hamming256(a0, a1, desc2 + 2 * (size_t)ib);
        if (d > TH_LOW) continue;
        if (!stereo1 && !stereo2) {
            const float dx = __fsub_rn(ex, k2.x), dy = __fsub_rn(ey, k2.y);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, sf2[k2.octave])) continue;
        }
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);
        if (den == 0) continue;
        const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
        if (!((double)dsqr < 3.84 * (double)sig2[k2.octave])) continue;
        key = min(key, ((unsigned)d << 20) | (0xfffffu - (unsigned)(r - it.b0)));
    }
    key = warp_min(key);
    if (lane == 0) match12[it.ia] = key == 0xffffffffu ? -1 : idx2[it.b0 + (int)(0xfffffu - (key & 0xfffffu))];
}

// One warp per map point (MapPoint::ComputeDistinctiveDescriptors R21/src/MapPoint.cc:242-307): all-pairs
// Hamming distances into scratch, then each row's median by counting (the k-th smallest is the value v with
// #{< v} <= k < #{<= v}), then the first row of least median.
__global__ void __launch_bounds__(128) distinctive_kernel(const uint4* __restrict__ desc, const int* __restrict__ ptr,
                                                          const long long* __restrict__ soff, int n_points,
                                                          unsigned short* __restrict__ scratch, int* __restrict__ best) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_points) return;
    const int first = ptr[p], N = ptr[p + 1] - first;
    if (N <= 0) { if (lane == 0) best[p] = -1; return; }
    unsigned short* D = scratch + soff[p];
    for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - i * N;
        const uint4 a0 = desc[2 * (size_t)(first + i)], a1 = desc[2 * (size_t)(first + i) + 1];
        D[idx] = (unsigned short)hamming256(a0, a1, desc + 2 * (size_t)(first + j));
    }
    __syncwarp();
    const int k = (int)(0.5 * (N - 1));
    unsigned key = 0xffffffffu;
    for (int i = lane; i < N; i += 32) {
        const unsigned short* row = D + (size_t)i * N;
        int median = 256;
        for (int j = 0; j < N; j++) {
            const int v = row[j];
            int lt = 0, le = 0;
            for (int l = 0; l < N; l++) { const int u = row[l]; lt += u < v; le += u <= v; }
            if (lt <= k && k < le) { median = v; break; }
        }
        key = min(key, ((unsigned)median << 20) | (unsigned)i);   // least median, first index on ties
    }
    key = warp_min(key);
    if (lane == 0) best[p] = (int)(key & 0xfffff);
}

struct StereoOut { float u_right, depth; int sad; int ok; };

// Frame::ComputeStereoMatches :504-628.  A CTA of 8 warps takes 32 left key points of one stereo pair (blockIdx.y = pair of a
// batch: the key point / descriptor arrays advance by `cap` entries per pair, the pyramids by one frame block, the per-pair
// counts come from cntL / cntR -- device arrays; nullptr for the single-pair call, which passes nl / nr).
//   * the right key points are reduced ONCE per CTA to what the candidate test needs -- x, the row band [minr, maxr] of
//     vRowIndices (:481-498) clamped to rows >= 0, the octave: 11 bytes each in shared memory, sorted by row bin -- instead of
//     every warp re-reading all 28-byte key points and redoing ceilf/floorf for each of its left key points (2.6 GB of L1/L2
//     reads per 64 pairs);
//   * a warp scans the bins around its row 32 entries at a time for each of its 4 left key points, collects the candidates
//     with a ballot, then computes their Hamming distances a candidate per lane (all descriptor loads in flight);
//   * the 11 x 11 left window and the 11 x 21 right strip of the SAD search (:548-598) are staged in shared memory once (12
//     byte loads per lane instead of 83), the 11 x 11 (shift, window row) sums are spread over the lanes.
constexpr int kStWarps = 16, kStPerWarp = 4, kStLeftPerCta = kStWarps * kStPerWarp, kStChunk = 2048, kStCand = 64, kStBins = 128;
__global__ void __launch_bounds__(kStWarps * 32, 2) stereo_kernel(const orb_keypoint_t* __restrict__ kl, const uint4* __restrict__ dl, int nl,
                                                     const orb_keypoint_t* __restrict__ kr, const uint4* __restrict__ dr, int nr,
                                                     const int32_t* __restrict__ cntL, const int32_t* __restrict__ cntR, int cap,
                                                     const float* __restrict__ sfs, const float* __restrict__ isfs,
                                                     const uint8_t* __restrict__ pyrL, const uint8_t* __restrict__ pyrR, size_t pyr_stride,
                                                     const uint8_t* __restrict__ l0L, const uint8_t* __restrict__ l0R, int l0_pitch,
                                                     size_t l0_strideL, size_t l0_strideR,
                                                     const LevelGeom* __restrict__ geom, int n_rows, float mbf, float max_d,
                                                     StereoOut* __restrict__ out) {
    __shared__ float s_x[kStChunk];
    __shared__ unsigned s_rows[kStChunk];                 // minr | (maxr + 1) << 16, both clamped to [0, 65535]
    __shared__ uint8_t s_oct[kStChunk];
    __shared__ unsigned short s_idx[kStChunk];           // position in the chunk of the key point staged in slot i
    __shared__ int s_hist[kStBins], s_start[kStBins], s_w;
    __shared__ int s_cand[kStWarps][kStCand];
    __shared__ int s_part[kStWarps][128];
    __shared__ uint8_t s_winL[kStWarps][128], s_winR[kStWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const size_t f = blockIdx.y;
        if (cntL) { nl = min(cntL[f], cap); nr = min(cntR[f], cap); }
        kl += f * cap; dl += 2 * f * cap; kr += f * cap; dr += 2 * f * cap; out += f * cap;
        pyrL += f * pyr_stride; pyrR += f * pyr_stride; l0L += f * l0_strideL; l0R += f * l0_strideR;
    }
    const int base = blockIdx.x * kStLeftPerCta;
    if (base >= nl) return;
    // ---- my four left key points
    int iLs[kStPerWarp], rowv[kStPerWarp], lev[kStPerWarp];
    float minUs[kStPerWarp], maxUs[kStPerWarp];
    bool alive[kStPerWarp];
    unsigned best[kStPerWarp];
#pragma unroll
    for (int j = 0; j < kStPerWarp; j++) {
        const int iL = base + j * kStWarps + warp;
        iLs[j] = iL; best[j] = 0xffffffffu; alive[j] = false; rowv[j] = 0; lev[j] = 0; minUs[j] = 0.f; maxUs[j] = 0.f;
        if (iL < nl) {
            const orb_keypoint_t kpL = kl[iL];
            rowv[j] = (int)kpL.y; lev[j] = kpL.octave;
            minUs[j] = __fsub_rn(kpL.x, max_d); maxUs[j] = kpL.x;
            alive[j] = rowv[j] >= 0 && rowv[j] < n_rows && !(maxUs[j] < 0);
        }
    }
    // ---- candidate search over the right key points, kStChunk at a time.  Staging sorts them by the first row of their band
    // (a counting sort over bins of `h` rows): iR is a candidate of `row` only if minr lies in (row - W, row], W = the widest
    // band staged, so a left key point scans the few bins that cover those rows -- ~7 % of the right key points -- and applies
    // the exact test there.  The order inside a bin is arbitrary: the winner is the least (distance, iR) key whatever the order.
    const int h = max(8, (n_rows + kStBins - 1) / kStBins);
    auto reduce_kp = [&](const orb_keypoint_t& kpR, unsigned& rw) {
        // membership of iR in vRowIndices[row] (:481-498): minr <= row <= maxr; row >= 0, so the clamps change nothing
        const float r = __fmul_rn(2.0f, sfs[kpR.octave]);
        const int maxr = (int)ceilf(__fadd_rn(kpR.y, r)), minr = (int)floorf(__fsub_rn(kpR.y, r));
        const int lo = min(max(minr, 0), 65535), hi = min(max(maxr + 1, 0), 65535);
        rw = (unsigned)lo | ((unsigned)hi << 16);
        return min(lo, n_rows - 1) / h;
    };
    for (int c0 = 0; c0 < nr; c0 += kStChunk) {
        const int nc = min(kStChunk, nr - c0);
        if (c0) __syncthreads();
        for (int i = threadIdx.x; i < kStBins; i += blockDim.x) s_hist[i] = 0;
        if (threadIdx.x == 0) s_w = 0;
        __syncthreads();
        int wmax = 0;
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            unsigned rw;
            const int bin = reduce_kp(kr[c0 + i], rw);
            atomicAdd(&s_hist[bin], 1);
            wmax = max(wmax, (int)(rw >> 16) - (int)(rw & 0xffffu));
        }
        wmax = __reduce_max_sync(0xffffffffu, wmax);
        if (lane == 0 && wmax > 0) atomicMax(&s_w, wmax);
        __syncthreads();
        if (warp == 0) {                                 // exclusive prefix of the bin counts -> first slot of every bin
            constexpr int kPer = kStBins / 32;
            int c[kPer], sum = 0;
#pragma unroll
            for (int k = 0; k < kPer; k++) { c[k] = s_hist[lane * kPer + k]; sum += c[k]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int at = incl - sum;
#pragma unroll
            for (int k = 0; k < kPer; k++) { s_start[lane * kPer + k] = at; s_hist[lane * kPer + k] = at; at += c[k]; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nc; i += blockDim.x) {
            const orb_keypoint_t kpR = kr[c0 + i];
            unsigned rw;
            const int bin = reduce_kp(kpR, rw);
            const int at = atomicAdd(&s_hist[bin], 1);   // afterwards s_hist[b] = end of bin b
            s_x[at] = kpR.x; s_rows[at] = rw; s_oct[at] = (uint8_t)min(max(kpR.octave, 0), 255); s_idx[at] = (unsigned short)i;
        }
        __syncthreads();
        const int W = max(s_w, 1);
#pragma unroll
        for (int j = 0; j < kStPerWarp; j++) {
            if (!alive[j]) continue;                     // warp-uniform
            const int row = rowv[j], lo_oct = lev[j] - 1, hi_oct = lev[j] + 1;
            const float minU = minUs[j], maxU = maxUs[j];
            const uint4 a0 = dl[2 * (size_t)iLs[j]], a1 = dl[2 * (size_t)iLs[j] + 1];
            unsigned b = best[j];
            int count = 0;
            const int i_begin = s_start[max(row - W + 1, 0) / h], i_end = s_hist[row / h];
            for (int i0 = i_begin; i0 < i_end; i0 += 32) {
                const int i = i0 + lane;
                bool cand = false;
                if (i < i_end) {
                    const unsigned rw = s_rows[i];
                    const int oc = s_oct[i];
                    const float x = s_x[i];
                    cand = row >= (int)(rw & 0xffffu) && row < (int)(rw >> 16) && oc >= lo_oct && oc <= hi_oct && x >= minU && x <= maxU;
                }
                const unsigned m = __ballot_sync(0xffffffffu, cand);
                if (m) {
                    const int pos = count + __popc(m & ((1u << lane) - 1u));
                    if (cand) {
                        const int iR = c0 + s_idx[i];
                        if (pos < kStCand) s_cand[warp][pos] = iR;
                        else {                            // more candidates than the list holds (crowded rows): compare here
                            const int d = hamming256(a0, a1, dr + 2 * (size_t)iR);
                            if (d < TH_HIGH) b = min(b, ((unsigned)d << 20) | (unsigned)iR);
                        }
                    }
                    count += __popc(m);
                }
            }
            __syncwarp();
            for (int c = lane; c < min(count, kStCand); c += 32) {
                const int iR = s_cand[warp][c];
                const int d = hamming256(a0, a1, dr + 2 * (size_t)iR);
                if (d < TH_HIGH) b = min(b, ((unsigned)d << 20) | (unsigned)iR);   // first strict minimum: least (distance, iR)
            }
            __syncwarp();
            best[j] = b;
        }
    }
    // ---- sub-pixel refinement by SAD of each surviving key point (:548-628)
#pragma unroll
    for (int j = 0; j < kStPerWarp; j++) {
        if (iLs[j] >= nl) continue;
        StereoOut o = {-1.f, -1.f, 0, 0};
        const unsigned bw = warp_min(best[j]);
        if (alive[j] && bw != 0xffffffffu && (int)(bw >> 20) < (TH_HIGH + TH_LOW) / 2) {
            const orb_keypoint_t kpL = kl[iLs[j]];
            const int levelL = kpL.octave;
            const float uL = kpL.x;
            const int bestIdxR = (int)(bw & 0xfffff);
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = isfs[levelL];
            const int cu = (int)roundf(__fmul_rn(kpL.x, scaleFactor));
            const int cv = (int)roundf(__fmul_rn(kpL.y, scaleFactor));
            const int cr0 = (int)roundf(__fmul_rn(uR0, scaleFactor));
            const LevelGeom g = geom[levelL];
            const int w = 5, L = 5;
            if (!(cr0 < 0 || cr0 + L + w + 1 >= g.w)) {   // iniu < 0 || endu >= cols  (:579-582)
                // level 0 is the input image itself, levels >= 1 live in the padded pyramid planes
                const int pitch = levelL ? g.pitch : l0_pitch;
                const uint8_t* PL = levelL ? pyrL + g.plane_off + (size_t)kEdge * g.pitch + kXPad : l0L;
                const uint8_t* PR = levelL ? pyrR + g.plane_off + (size_t)kEdge * g.pitch + kXPad : l0R;
                PL += (ptrdiff_t)(cv - w) * pitch + (cu - w);
                PR += (ptrdiff_t)(cv - w) * pitch + (cr0 - L - w);
                uint8_t* wl = s_winL[warp];
                uint8_t* wr = s_winR[warp];
                {   // lanes 0..20 copy a column of the right strip each, lanes 21..31 a column of the left window
                    const bool right = lane < 21;
                    const uint8_t* src = right ? PR + lane : PL + (lane - 21);
                    uint8_t* dst = right ? wr + lane : wl + (lane - 21);
                    const int dstride = right ? 21 : 11;
#pragma unroll
                    for (int dy = 0; dy < 11; dy++) dst[dy * dstride] = src[(ptrdiff_t)dy * pitch];
                }
                __syncwarp();
                const int cL = wl[5 * 11 + 5];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int t = lane + 32 * k;          // t = shift * 11 + window row
                    if (t < 121) {
                        const int sft = t / 11, dy = t - 11 * sft;
                        const int c = cL - (int)wr[5 * 21 + sft + 5];
                        const uint8_t* a = wl + dy * 11;
                        const uint8_t* bb = wr + dy * 21 + sft;
                        int acc = 0;
#pragma unroll
                        for (int i = 0; i < 11; i++) acc += abs((int)a[i] - (int)bb[i] - c);
                        s_part[warp][t] = acc;
                    }
                }
                __syncwarp();
                int sad = 0;
                if (lane < 11) {
#pragma unroll
                    for (int i = 0; i < 11; i++) sad += s_part[warp][lane * 11 + i];
                }
                const unsigned key = lane < 11 ? ((unsigned)sad << 4) | (unsigned)lane : 0xffffffffu;
                const unsigned bk = warp_min(key);        // least SAD, first shift on ties (:590-594)
                const int bestSad = (int)(bk >> 4), bs = (int)(bk & 15u), bestinc = bs - L;
                const float d1 = (float)__shfl_sync(0xffffffffu, sad, max(bs - 1, 0));
                const float d3 = (float)__shfl_sync(0xffffffffu, sad, min(bs + 1, 10));
                const float d2 = (float)bestSad;
                if (bestinc != -L && bestinc != L) {
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(sfs[levelL], __fadd_rn(__fadd_rn((float)cr0, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0 && disparity < max_d) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                            o.depth = __fdiv_rn(mbf, disparity);
                            o.u_right = bestuR;
                            o.sad = bestSad;
                            o.ok = 1;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) out[iLs[j]] = o;
    }
}


// Outlier cut of Frame::ComputeStereoMatches (:630-644) for a batch, one CTA per stereo pair: the median is the
// element of rank size/2 of the (SAD, index)-sorted list of accepted matches, every match with SAD >= 1.5*1.4*median
// is dropped (the reference walks the sorted list from the back and stops at the first SAD below the limit).  Only the
// SAD VALUE of that rank matters: it is the largest v with count(SAD < v) <= rank, found bit by bit (a SAD is at most
// 121 * 510 < 2^16) -- 16 counting rounds over the pair's results instead of a rank count per element (n^2 loads).
// Writes mvuRight / mvDepth (-1 where no match) and the kept count.
__global__ void __launch_bounds__(256) stereo_filter_kernel(const StereoOut* __restrict__ res, const int32_t* __restrict__ cntL, int cap,
                                                            float* __restrict__ u_right, float* __restrict__ depth,
                                                            int32_t* __restrict__ n_matches) {
    constexpr int kKeep = 4096;
    const size_t f = blockIdx.x;
    res += f * cap; u_right += f * cap; depth += f * cap;
    const int n = min(cntL[f], cap);
    __shared__ int s_sad[kKeep];                         // SAD of the accepted matches, 0x7fffffff otherwise
    __shared__ int s_cnt[18], s_kept;
    if (threadIdx.x < 18) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    auto sad_of = [&](int i) { return i < kKeep ? s_sad[i] : (res[i].ok ? res[i].sad : 0x7fffffff); };
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const StereoOut o = res[i];
        if (i < kKeep) s_sad[i] = o.ok ? o.sad : 0x7fffffff;
        mine += o.ok;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt[16], mine);
    __syncthreads();
    const int n_ok = s_cnt[16];
    int median = -1;
    if (n_ok > 0) {
        const int k = n_ok / 2;
        int v = 0;
        for (int bit = 15; bit >= 0; bit--) {
            const int t = v | (1 << bit);
            int below = 0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) below += sad_of(i) < t;
            below = __reduce_add_sync(0xffffffffu, below);
            if ((threadIdx.x & 31) == 0 && below) atomicAdd(&s_cnt[bit], below);
            __syncthreads();
            if (s_cnt[bit] <= k) v = t;
        }
        median = v;
    }
    const float th_dist = 1.5f * 1.4f * (float)median;
    int kept = 0;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        float u = -1.f, d = -1.f;
        if (i < n) {
            const StereoOut o = res[i];
            if (o.ok && (float)o.sad < th_dist) { u = o.u_right; d = o.depth; kept++; }
        }
        u_right[i] = u; depth[i] = d;
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (threadIdx.x == 0) n_matches[f] = s_kept;
}

// ---- host helpers -----------------------------------------------------------------------------------
static void shared_nodes(const orbm_featvec_t* a, const orbm_featvec_t* b, std::vector<NodePair>& out) {
    int i = 0, j = 0;   // merge walk of two sorted maps (R21 ORBmatcher.cc:175-263)
    while (i < a->n_nodes && j < b->n_nodes) {
        if (a->node_ids[i] == b->node_ids[j]) {
            out.push_back(NodePair{a->ptr[i], a->ptr[i + 1], b->ptr[j], b->ptr[j + 1]});
            i++; j++;
        } else if (a->node_ids[i] < b->node_ids[j]) i++;
        else j++;
    }
}

// ComputeThreeMaxima R21 :1601-1642 on bin counts
static void three_maxima(const int* cnt, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
static inline int rot_bin(float a1, float a2) {   // R21 :236-243 (factor = 1.0f/HISTO_LENGTH as written there)
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

static int run_bow(int mode, const uint8_t* dA, const uint8_t* vA, int nA, const orbm_featvec_t* fvA, const uint8_t* dB,
                   const uint8_t* vB, int nB, const orbm_featvec_t* fvB, float ratio, std::vector<int>& matchB,
                   std::vector<int>& matchA, int device) {
    if (!dA || !vA || !fvA || !dB || !fvB || nA < 0 || nB < 0 || (mode == 1 && !vB)) { set_error("SearchByBoW: bad arguments"); return ORB_ERR_ARG; }
    if (nB >= (1 << 20)) { set_error("SearchByBoW: too many features"); return ORB_ERR_ARG; }
    matchB.assign(nB, -1);
    matchA.assign(nA, -1);
    std::vector<NodePair> pairs;
    shared_nodes(fvA, fvB, pairs);
    if (pairs.empty() || nA == 0 || nB == 0) return ORB_OK;
    const int nia = fvA->ptr[fvA->n_nodes], nib = fvB->ptr[fvB->n_nodes];
    const size_t bytes = (size_t)nA * 32 + nA + (size_t)nia * 4 + (size_t)nB * 32 + nB + (size_t)nib * 4 +
                         pairs.size() * sizeof(NodePair) + (size_t)(nA + nB) * 4;
    MatchCtx& cx = match_ctx();
    if (!cx.begin(device, bytes + 16 * 256, bytes + (size_t)(nA + nB) * 4 + 24 * 256)) return ORB_ERR_CUDA;
    const uint4* d_dA = (const uint4*)cx.upload(dA, (size_t)nA * 32);
    const uint8_t* d_vA = (const uint8_t*)cx.upload(vA, nA);
    const int* d_iA = (const int*)cx.upload(fvA->idx, (size_t)nia * 4);
    const uint4* d_dB = (const uint4*)cx.upload(dB, (size_t)nB * 32);
    const uint8_t* d_vB = (const uint8_t*)cx.upload(vB ? vB : vA, vB ? nB : 0);
    const int* d_iB = (const int*)cx.upload(fvB->idx, (size_t)nib * 4);
    const NodePair* d_pairs = (const NodePair*)cx.upload(pairs.data(), pairs.size() * sizeof(NodePair));
    int* d_mB = (int*)cx.dalloc((size_t)nB * 4);
    int* d_mA = (int*)cx.dalloc((size_t)nA * 4);
    if (!d_dA || !d_vA || !d_iA || !d_dB || !d_vB || !d_iB || !d_pairs || !d_mB || !d_mA) return ORB_ERR_CUDA;
    ORB_CUDA_TRY(cudaMemsetAsync(d_mB, 0xff, (size_t)nB * 4, cx.s()));
    ORB_CUDA_TRY(cudaMemsetAsync(d_mA, 0xff, (size_t)nA * 4, cx.s()));
    const int np = (int)pairs.size();
    bow_kernel<<<(np * 32 + 127) / 128, 128, 0, cx.s()>>>(d_dA, d_vA, d_iA, d_dB, d_vB, d_iB, d_pairs, np, ratio, mode, d_mB, d_mA);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(matchB.data(), d_mB, (size_t)nB * 4) || !cx.download(matchA.data(), d_mA, (size_t)nA * 4) || !cx.finish())
        return ORB_ERR_CUDA;
    return ORB_OK;
}

}  // namespace orbcuda

using namespace orbcuda;

// accessor implemented in extractor.cu
extern "C" int orbx_internal_view(orbx_handle_t h, const uint8_t** d_pyr, const orbcuda::LevelGeom** d_geom,
                                  const orbcuda::LevelGeom** h_geom, orbcuda::FrameLayout* fl, int* device,
                                  const float** sf, const float** isf, const uint8_t** d_level0, int* level0_pitch);
extern "C" int orbx_internal_view_batch(orbx_handle_t h, const uint8_t** d_pyr, const orbcuda::LevelGeom** d_geom,
                                        const orbcuda::LevelGeom** h_geom, orbcuda::FrameLayout* fl, int* device,
                                        const float** d_tables, const uint8_t** d_level0, int* level0_pitch,
                                        size_t* level0_stride, int* n_frames, void* consumer);

extern "C" {

int orbm_search_by_bow_kf_f(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* kf_valid, int n_kf,
                            const orbm_featvec_t* fv_kf, const uint8_t* desc_f, const float* angle_f, int n_f,
                            const orbm_featvec_t* fv_f, float nnratio, int check_ori, int32_t* out_match_f, int* n_matches,
                            int device) {
    if (!out_match_f || !n_matches || (check_ori && (!angle_kf || !angle_f))) { set_error("SearchByBoW: bad arguments"); return ORB_ERR_ARG; }
    std::vector<int> mB, mA;
    const int rc = run_bow(0, desc_kf, kf_valid, n_kf, fv_kf, desc_f, nullptr, n_f, fv_f, nnratio, mB, mA, device);
    if (rc) return rc;
    int nmatches = 0;
    for (int j = 0; j < n_f; j++) nmatches += mB[j] >= 0;
    if (check_ori) {   // R21 :236-285
        int cnt[HISTO_LENGTH] = {0};
        std::vector<int> bin(n_f, -1);
        for (int j = 0; j < n_f; j++)
            if (mB[j] >= 0) { bin[j] = rot_bin(angle_kf[mB[j]], angle_f[j]); cnt[bin[j]]++; }
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(cnt, HISTO_LENGTH, ind1, ind2, ind3);
        for (int j = 0; j < n_f; j++)
            if (bin[j] >= 0 && bin[j] != ind1 && bin[j] != ind2 && bin[j] != ind3) { mB[j] = -1; nmatches--; }
    }
    for (int j = 0; j < n_f; j++) out_match_f[j] = mB[j];
    *n_matches = nmatches;
    return ORB_OK;
}

int orbm_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                             const orbm_featvec_t* fv1, const uint8_t* desc2, const float* angle2, const uint8_t* valid2,
                             int n2, const orbm_featvec_t* fv2, float nnratio, int check_ori, int32_t* out_match12,
                             int* n_matches, int device) {
    if (!out_match12 || !n_matches || (check_ori && (!angle1 || !angle2))) { set_error("SearchByBoW: bad arguments"); return ORB_ERR_ARG; }
    std::vector<int> mB, mA;
    const int rc = run_bow(1, desc1, valid1, n1, fv1, desc2, valid2, n2, fv2, nnratio, mB, mA, device);
    if (rc) return rc;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) nmatches += mA[i] >= 0;
    if (check_ori) {   // R21 :607-653
        int cnt[HISTO_LENGTH] = {0};
        std::vector<int> bin(n1, -1);
        for (int i = 0; i < n1; i++)
            if (mA[i] >= 0) { bin[i] = rot_bin(angle1[i], angle2[mA[i]]); cnt[bin[i]]++; }
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(cnt, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < n1; i++)
            if (bin[i] >= 0 && bin[i] != ind1 && bin[i] != ind2 && bin[i] != ind3) { mA[i] = -1; nmatches--; }
    }
    for (int i = 0; i < n1; i++) out_match12[i] = mA[i];
    *n_matches = nmatches;
    return ORB_OK;
}

int orbm_search_for_triangulation(const uint8_t* desc1, const orbm_tri_feature_t* f1, int n1, const orbm_featvec_t* fv1,
                                  const uint8_t* desc2, const orbm_tri_feature_t* f2, int n2, const orbm_featvec_t* fv2,
                                  const float* F12, float ex, float ey, const float* scale_factors2,
                                  const float* level_sigma2_2, int only_stereo, int check_ori, int32_t* out_pairs,
                                  int cap_pairs, int* n_matches, int device) {
    if (!desc1 || !f1 || !fv1 || !desc2 || !f2 || !fv2 || !F12 || !scale_factors2 || !level_sigma2_2 || !n_matches || n1 < 0 || n2 < 0 ||
        (cap_pairs > 0 && !out_pairs)) { set_error("SearchForTriangulation: bad arguments"); return ORB_ERR_ARG; }
    *n_matches = 0;
    std::vector<NodePair> pairs;
    shared_nodes(fv1, fv2, pairs);
    std::vector<TriItem> items;
    for (const NodePair& p : pairs)
        for (int q = p.a0; q < p.a1; q++) {
            const int ia = fv1->idx[q];
            if (f1[ia].has_mp) continue;                         // :700-702
            if (only_stereo && !(f1[ia].u_right >= 0)) continue;  // :704-708
            if (p.b1 - p.b0 >= (1 << 20)) { set_error("node too large"); return ORB_ERR_ARG; }
            items.push_back(TriItem{ia, p.b0, p.b1});
        }
    std::vector<int> m12(n1, -1);
    if (!items.empty() && n2 > 0) {
        int max_oct = 0;
        for (int i = 0; i < n2; i++) max_oct = std::max(max_oct, f2[i].octave);
        const int ni2 = fv2->ptr[fv2->n_nodes];
        const size_t bytes = (size_t)n1 * (32 + sizeof(TriFeat) + 4) + (size_t)n2 * (32 + sizeof(TriFeat)) + (size_t)ni2 * 4 +
                             items.size() * sizeof(TriItem) + 36 + (size_t)(max_oct + 1) * 8;
        MatchCtx& cx = match_ctx();
        if (!cx.begin(device, bytes + 16 * 256, bytes + (size_t)n1 * 4 + 24 * 256)) return ORB_ERR_CUDA;
        const uint4* d_d1 = (const uint4*)cx.upload(desc1, (size_t)n1 * 32);
        const TriFeat* d_f1 = (const TriFeat*)cx.upload(f1, (size_t)n1 * sizeof(TriFeat));
        const uint4* d_d2 = (const uint4*)cx.upload(desc2, (size_t)n2 * 32);
        const TriFeat* d_f2 = (const TriFeat*)cx.upload(f2, (size_t)n2 * sizeof(TriFeat));
        const int* d_i2 = (const int*)cx.upload(fv2->idx, (size_t)ni2 * 4);
        const TriItem* d_it = (const TriItem*)cx.upload(items.data(), items.size() * sizeof(TriItem));
        const float* d_F = (const float*)cx.upload(F12, 36);
        const float* d_sf = (const float*)cx.upload(scale_factors2, (size_t)(max_oct + 1) * 4);
        const float* d_s2 = (const float*)cx.upload(level_sigma2_2, (size_t)(max_oct + 1) * 4);
        int* d_m = (int*)cx.dalloc((size_t)n1 * 4);
        if (!d_d1 || !d_f1 || !d_d2 || !d_f2 || !d_i2 || !d_it || !d_F || !d_sf || !d_s2 || !d_m) return ORB_ERR_CUDA;
        ORB_CUDA_TRY(cudaMemsetAsync(d_m, 0xff, (size_t)n1 * 4, cx.s()));
        const int ni = (int)items.size();
        triangulation_kernel<<<(ni * 32 + 127) / 128, 128, 0, cx.s()>>>(d_d1, d_f1, d_d2, d_f2, d_i2, d_it, ni, d_F, ex, ey, d_sf, d_s2,
                                                                         only_stereo, d_m);
        ORB_CUDA_TRY(cudaGetLastError());
        if (!cx.download(m12.data(), d_m, (size_t)n1 * 4) || !cx.finish()) return ORB_ERR_CUDA;
    }
    int nmatches = 0;
    for (int i = 0; i < n1; i++) nmatches += m12[i] >= 0;
    if (check_ori) {   // :775-808
        int cnt[HISTO_LENGTH] = {0};
        std::vector<int> bin(n1, -1);
        for (int i = 0; i < n1; i++)
            if (m12[i] >= 0) { bin[i] = rot_bin(f1[i].angle, f2[m12[i]].angle); cnt[bin[i]]++; }
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(cnt, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < n1; i++)
            if (bin[i] >= 0 && bin[i] != ind1 && bin[i] != ind2 && bin[i] != ind3) { m12[i] = -1; nmatches--; }
    }
    int np = 0;
    for (int i = 0; i < n1; i++) {   // :812-820
        if (m12[i] < 0) continue;
        if (np < cap_pairs) { out_pairs[2 * np] = i; out_pairs[2 * np + 1] = m12[i]; }
        np++;
    }
    *n_matches = nmatches;
    return np > cap_pairs ? ORB_ERR_CAPACITY : ORB_OK;
}

int orbm_distinctive_descriptors(const uint8_t* desc, const int32_t* ptr, int n_points, int32_t* best, int device) {
    if (n_points < 0 || !ptr || !best || (n_points > 0 && ptr[n_points] > 0 && !desc)) { set_error("orbm_distinctive_descriptors: bad arguments"); return ORB_ERR_ARG; }
    if (n_points == 0) return ORB_OK;
    std::vector<long long> soff(n_points + 1, 0);
    for (int p = 0; p < n_points; p++) {
        const long long N = ptr[p + 1] - ptr[p];
        if (N < 0 || N >= (1 << 20)) { set_error("orbm_distinctive_descriptors: bad CSR"); return ORB_ERR_ARG; }
        soff[p + 1] = soff[p] + N * N;
    }
    const size_t up = (size_t)ptr[n_points] * 32 + (size_t)(n_points + 1) * 12;
    MatchCtx& cx = match_ctx();
    if (!cx.begin(device, up + (size_t)soff[n_points] * 2 + (size_t)n_points * 4 + 16 * 256, up + (size_t)n_points * 4 + 16 * 256)) return ORB_ERR_CUDA;
    const uint4* d_desc = (const uint4*)cx.upload(desc, (size_t)ptr[n_points] * 32);
    const int* d_ptr = (const int*)cx.upload(ptr, (size_t)(n_points + 1) * 4);
    const long long* d_soff = (const long long*)cx.upload(soff.data(), (size_t)(n_points + 1) * 8);
    unsigned short* d_scr = (unsigned short*)cx.dalloc((size_t)soff[n_points] * 2);
    int* d_best = (int*)cx.dalloc((size_t)n_points * 4);
    if (!d_desc || !d_ptr || !d_soff || !d_scr || !d_best) return ORB_ERR_CUDA;
    distinctive_kernel<<<(n_points * 32 + 127) / 128, 128, 0, cx.s()>>>(d_desc, d_ptr, d_soff, n_points, d_scr, d_best);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(best, d_best, (size_t)n_points * 4) || !cx.finish()) return ORB_ERR_CUDA;
    return ORB_OK;
}

int orbm_stereo_matches(orbx_handle_t hl, orbx_handle_t hr, const orb_keypoint_t* keys_left, const uint8_t* desc_left,
                        int n_left, const orb_keypoint_t* keys_right, const uint8_t* desc_right, int n_right, float mbf,
                        float mb, float* u_right, float* depth, int* n_matches) {
    if (!hl || !hr || n_left < 0 || n_right < 0 || !u_right || !depth || (n_left && (!keys_left || !desc_left)) ||
        (n_right && (!keys_right || !desc_right))) { set_error("orbm_stereo_matches: bad arguments"); return ORB_ERR_ARG; }
    for (int i = 0; i < n_left; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }   // :473-474
    if (n_matches) *n_matches = 0;
    if (n_left == 0 || n_right == 0) return ORB_OK;
    const uint8_t *pl, *pr; const LevelGeom *dgl, *dgr, *hgl, *hgr; FrameLayout fll, flr; int devl, devr;
    const float *sf, *isf, *sfr, *isfr;
    const uint8_t *l0l, *l0r; int l0pl, l0pr;
    if (orbx_internal_view(hl, &pl, &dgl, &hgl, &fll, &devl, &sf, &isf, &l0l, &l0pl) ||
        orbx_internal_view(hr, &pr, &dgr, &hgr, &flr, &devr, &sfr, &isfr, &l0r, &l0pr) ||
        !pl || !pr || !l0l || !l0r || l0pl != l0pr || devl != devr || fll.width != flr.width || fll.height != flr.height || fll.nlevels != flr.nlevels) {
        set_error("orbm_stereo_matches: the two extractors must have processed same-size images on one device");
        return ORB_ERR_ARG;
    }
    if (n_right >= (1 << 20) || hgl[0].h > 65535) return ORB_ERR_ARG;     // index / row-band packing of stereo_kernel
    std::vector<StereoOut> res(n_left);
    {
        const float max_d = mbf / mb;   // maxD = mbf/minZ, minZ = mb  (:501-503)
        const size_t bytes = (size_t)(n_left + n_right) * (sizeof(orb_keypoint_t) + 32) + (size_t)fll.nlevels * 8;
        MatchCtx& cx = match_ctx();
        if (!cx.begin(devl, bytes + (size_t)n_left * sizeof(StereoOut) + 16 * 256, bytes + (size_t)n_left * sizeof(StereoOut) + 16 * 256)) return ORB_ERR_CUDA;
        const orb_keypoint_t* d_kl = (const orb_keypoint_t*)cx.upload(keys_left, (size_t)n_left * sizeof(orb_keypoint_t));
        const uint4* d_dl = (const uint4*)cx.upload(desc_left, (size_t)n_left * 32);
        const orb_keypoint_t* d_kr = (const orb_keypoint_t*)cx.upload(keys_right, (size_t)n_right * sizeof(orb_keypoint_t));
        const uint4* d_dr = (const uint4*)cx.upload(desc_right, (size_t)n_right * 32);
        const float* d_sf = (const float*)cx.upload(sf, (size_t)fll.nlevels * 4);
        const float* d_isf = (const float*)cx.upload(isf, (size_t)fll.nlevels * 4);
        StereoOut* d_out = (StereoOut*)cx.dalloc((size_t)n_left * sizeof(StereoOut));
        if (!d_kl || !d_dl || !d_kr || !d_dr || !d_sf || !d_isf || !d_out) return ORB_ERR_CUDA;
        stereo_kernel<<<(n_left + kStLeftPerCta - 1) / kStLeftPerCta, kStWarps * 32, 0, cx.s()>>>(d_kl, d_dl, n_left, d_kr, d_dr, n_right, nullptr, nullptr, 0, d_sf, d_isf,
                                                                      pl, pr, 0, l0l, l0r, l0pl, 0, 0, dgl, hgl[0].h, mbf, max_d, d_out);
        ORB_CUDA_TRY(cudaGetLastError());
        if (!cx.download(res.data(), d_out, (size_t)n_left * sizeof(StereoOut)) || !cx.finish()) return ORB_ERR_CUDA;
    }
    std::vector<std::pair<int, int>> distIdx;
    for (int i = 0; i < n_left; i++)
        if (res[i].ok) { u_right[i] = res[i].u_right; depth[i] = res[i].depth; distIdx.push_back(std::make_pair(res[i].sad, i)); }
    int kept = (int)distIdx.size();
    if (!distIdx.empty()) {   // :630-644 (the reference indexes an empty vector here; guarded)
        std::sort(distIdx.begin(), distIdx.end());
        const float median = (float)distIdx[distIdx.size() / 2].first;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
            if (distIdx[i].first < thDist) break;
            u_right[distIdx[i].second] = -1; depth[distIdx[i].second] = -1; kept--;
        }
    }
    if (n_matches) *n_matches = kept;
    return ORB_OK;
}

int orbm_stereo_matches_batch_device(orbx_handle_t hl, orbx_handle_t hr, const void* d_keys_left, const uint8_t* d_desc_left,
                                     const int32_t* d_counts_left, const void* d_keys_right, const uint8_t* d_desc_right,
                                     const int32_t* d_counts_right, int n_pairs, int cap, float mbf, float mb, void* d_scratch,
                                     float* d_u_right, float* d_depth, int32_t* d_n_matches, void* stream) {
    if (!hl || !hr || n_pairs < 1 || cap < 1 || cap >= (1 << 20) || !d_keys_left || !d_desc_left || !d_counts_left || !d_keys_right ||
        !d_desc_right || !d_counts_right || !d_scratch || !d_u_right || !d_depth || !d_n_matches) {
        set_error("orbm_stereo_matches_batch_device: bad arguments");
        return ORB_ERR_ARG;
    }
    const uint8_t *pl, *pr; const LevelGeom *dgl, *dgr, *hgl, *hgr; FrameLayout fll, flr; int devl, devr, nfl, nfr;
    const float *d_sfl, *d_sfr; const uint8_t *l0l, *l0r; int l0pl, l0pr; size_t l0sl, l0sr;
    if (orbx_internal_view_batch(hl, &pl, &dgl, &hgl, &fll, &devl, &d_sfl, &l0l, &l0pl, &l0sl, &nfl, stream) ||
        orbx_internal_view_batch(hr, &pr, &dgr, &hgr, &flr, &devr, &d_sfr, &l0r, &l0pr, &l0sr, &nfr, stream) ||
        !pl || !pr || !l0l || !l0r || l0pl != l0pr || devl != devr || fll.width != flr.width || fll.height != flr.height ||
        fll.nlevels != flr.nlevels || nfl < n_pairs || nfr < n_pairs || hgl[0].h > 65535) {
        set_error("orbm_stereo_matches_batch_device: the two extractors must hold a batch of >= n_pairs same-size images on one device");
        return ORB_ERR_ARG;
    }
    ORB_CUDA_TRY(cudaSetDevice(devl));
    cudaStream_t s = (cudaStream_t)stream;
    StereoOut* d_res = (StereoOut*)d_scratch;
    const float max_d = mbf / mb;   // maxD = mbf/minZ, minZ = mb  (:501-503)
    stereo_kernel<<<dim3((cap + kStLeftPerCta - 1) / kStLeftPerCta, n_pairs), kStWarps * 32, 0, s>>>(
        (const orb_keypoint_t*)d_keys_left, (const uint4*)d_desc_left, 0, (const orb_keypoint_t*)d_keys_right, (const uint4*)d_desc_right, 0,
        d_counts_left, d_counts_right, cap, d_sfl, d_sfl + kMaxLevels, pl, pr, (size_t)fll.pyr_bytes, l0l, l0r, l0pl, l0sl, l0sr, dgl, hgl[0].h,
        mbf, max_d, d_res);
    stereo_filter_kernel<<<n_pairs, 256, 0, s>>>(d_res, d_counts_left, cap, d_u_right, d_depth, d_n_matches);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

size_t orbm_stereo_scratch_bytes(int n_pairs, int cap) { return (size_t)std::max(n_pairs, 0) * (size_t)std::max(cap, 0) * sizeof(StereoOut); }

}  // extern "C"

