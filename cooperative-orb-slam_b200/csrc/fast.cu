// fast.cu -- K3: FAST-9/16 detection with the reference's per-cell semantics
// (replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree, R21/src/ORBextractor.cc:765-829,
// i.e. cv::FAST(cell, iniThFAST, nms) with the minThFAST retry on empty cells).
//
// Two kernels:
//   fast_score_kernel  dense per-pixel corner score at minThFAST over [19,W-19)x[19,H-19) of every
//                      level (the union of all cell interiors).  score = cornerScore<16>() =
//                      (max over the 16 arcs of 9 of max(min(I_p - I_k), min(I_k - I_p))) - 1 where that
//                      maximum exceeds the threshold, else 0.  FAST(th=20) == {FAST(th=7): score >= 20}
//                       so one map serves both thresholds.
//   fast_nms_kernel    dense strict 3x3 non-max suppression that ignores neighbours outside the pixel's
//                      cell interior (cv::FAST is called on the cell sub-image); survivors are appended
//                      to the level's candidate list.  The per-cell iniTh -> minTh fallback happens in
//                      the quadtree kernel's gather (octree.cu).
#include "internal.h"
#include <cstdlib>

namespace orbcuda {

constexpr int kFastFmaDefault = 8;      // two-input min/max per side on the FMA pipe (of 24)
constexpr int kFastRows = 14;   // rows per strip (two 7-row register rotations)

__device__ __forceinline__ uint32_t fun16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); }
__device__ __forceinline__ uint32_t mn3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t mx3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }

// 2-input min/max of non-negative pairs on the FMA pipe (the integer min/max all issue on the half-rate ALU pipe, which
// bounds this kernel): with a pixel p stored as the fp16 value 1024 + p (bit pattern 0x6400 + p -- ordered and spaced
// like the integer, so the 16-bit integer min/max/subtract work on it unchanged), max(x, y) = y + relu(x - y) and
// min(x, y) = y - relu(y - x) are exact in fp16 (|x - y| <= 255).
__device__ __forceinline__ uint32_t fmax2(uint32_t x, uint32_t y) {
    uint32_t t, r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(t) : "r"(y), "r"(0xBC00BC00u), "r"(x));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(y), "r"(t));
    return r;
}
__device__ __forceinline__ uint32_t fmin2(uint32_t x, uint32_t y) {
    uint32_t t, r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(t) : "r"(x), "r"(0xBC00BC00u), "r"(y));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(y), "r"(t));
    return r;
}
template <bool HF> __device__ __forceinline__ uint32_t mx2(uint32_t a, uint32_t b) { return HF ? fmax2(a, b) : __vmaxs2(a, b); }
template <bool HF> __device__ __forceinline__ uint32_t mn2(uint32_t a, uint32_t b) { return HF ? fmin2(a, b) : __vmins2(a, b); }

// One image row around a 4-pixel group at x0 as 16-bit pairs:
//   a0=(p[x0-4],p[x0-2]) a1=(p[x0-3],p[x0-1]) b0=(p[x0],p[x0+2]) b1=(p[x0+1],p[x0+3])
//   c0=(p[x0+4],p[x0+6]) c1=(p[x0+5],p[x0+7])
struct Row6 { uint32_t a0, a1, b0, b1, c0, c1; };

template <bool HF> __device__ __forceinline__ Row6 unpack_row6(uint32_t w0, uint32_t w1, uint32_t w2);
template <bool HF> __device__ __forceinline__ Row6 load_row6(const uint8_t* p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    return unpack_row6<HF>(w[0], w[1], w[2]);
}
template <bool HF> __device__ __forceinline__ Row6 unpack_row6(uint32_t w0, uint32_t w1, uint32_t w2) {
    // one PRMT per pair: bytes (0, 2) resp. (1, 3) of the word into the low bytes of the two 16-bit lanes, the high bytes
    // from B (HF: 0x64, every lane then holds the fp16 value 1024 + p)
    const uint32_t B = HF ? 0x64646464u : 0u;
    Row6 r;
    r.a0 = __byte_perm(w0, B, 0x4240); r.a1 = __byte_perm(w0, B, 0x4341);
    r.b0 = __byte_perm(w1, B, 0x4240); r.b1 = __byte_perm(w1, B, 0x4341);
    r.c0 = __byte_perm(w2, B, 0x4240); r.c1 = __byte_perm(w2, B, 0x4341);
    return r;
}
// E<t>(row) = (p[x0+t], p[x0+t+2])
__device__ __forceinline__ uint32_t Em3(const Row6& r) { return r.a1; }
__device__ __forceinline__ uint32_t Em2(const Row6& r) { return fun16(r.a0, r.b0); }
__device__ __forceinline__ uint32_t Em1(const Row6& r) { return fun16(r.a1, r.b1); }
__device__ __forceinline__ uint32_t E0(const Row6& r) { return r.b0; }
__device__ __forceinline__ uint32_t E1(const Row6& r) { return r.b1; }
__device__ __forceinline__ uint32_t E2(const Row6& r) { return fun16(r.b0, r.c0); }
__device__ __forceinline__ uint32_t E3(const Row6& r) { return fun16(r.b1, r.c1); }
__device__ __forceinline__ uint32_t E4(const Row6& r) { return r.c0; }

// max over the 16 circular 9-arcs of min(arc), and min over them of max(arc), on two pixels at once.
// Arcs k and k+1 (k even) share the eight pixels k+1..k+8: max(min(p[k], X), min(X, p[k+9])) = min(X, max(p[k], p[k+9])) with
// X = min(p[k+1..k+8]) = min of two 4-runs that start at odd positions.  Per side that is 8 pair minima (odd starts), 8 4-run
// minima, 8 outer maxima, 8 three-input minima and a 4-instruction tree: 36 operations as with the 3-runs formulation this
// replaces, but 25 of them are TWO-input -- and those can be issued on either half-rate pipe (ALU: one VIMNMX; FMA: fma.relu +
// add on the fp16 view).  The kernel is bound by the ALU pipe, so the split (FA of the 24 flexible operations per side go to the
// FMA pipe) is what balances the two.
template <bool HF, bool FMA_PIPE> __device__ __forceinline__ uint32_t mx2p(uint32_t a, uint32_t b) { return (HF && FMA_PIPE) ? fmax2(a, b) : __vmaxs2(a, b); }
template <bool HF, bool FMA_PIPE> __device__ __forceinline__ uint32_t mn2p(uint32_t a, uint32_t b) { return (HF && FMA_PIPE) ? fmin2(a, b) : __vmins2(a, b); }
template <bool HF, int FA> __device__ __forceinline__ void arc_extrema(const uint32_t (&p)[16], uint32_t& max_of_min, uint32_t& min_of_max) {
    // flexible operation number f (0..23 per side: pair minima 0..7, outer maxima 8..15, 4-run minima 16..23) runs on the FMA pipe if f < FA
    uint32_t lo2[8], hi2[8], lo4[8], hi4[8], ox[8], on[8], ra[8], rb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo2[i] = (i < FA) ? mn2p<HF, true>(p[2 * i + 1], p[(2 * i + 2) & 15]) : mn2p<HF, false>(p[2 * i + 1], p[(2 * i + 2) & 15]);
        hi2[i] = (i < FA) ? mx2p<HF, true>(p[2 * i + 1], p[(2 * i + 2) & 15]) : mx2p<HF, false>(p[2 * i + 1], p[(2 * i + 2) & 15]);
        ox[i] = (8 + i < FA) ? mx2p<HF, true>(p[2 * i], p[(2 * i + 9) & 15]) : mx2p<HF, false>(p[2 * i], p[(2 * i + 9) & 15]);
        on[i] = (8 + i < FA) ? mn2p<HF, true>(p[2 * i], p[(2 * i + 9) & 15]) : mn2p<HF, false>(p[2 * i], p[(2 * i + 9) & 15]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo4[i] = (16 + i < FA) ? mn2p<HF, true>(lo2[i], lo2[(i + 1) & 7]) : mn2p<HF, false>(lo2[i], lo2[(i + 1) & 7]);
        hi4[i] = (16 + i < FA) ? mx2p<HF, true>(hi2[i], hi2[(i + 1) & 7]) : mx2p<HF, false>(hi2[i], hi2[(i + 1) & 7]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ra[i] = mn3(lo4[i], lo4[(i + 2) & 7], ox[i]);      // max(arc 2i, arc 2i+1) of the minima
        rb[i] = mx3(hi4[i], hi4[(i + 2) & 7], on[i]);      // min(arc 2i, arc 2i+1) of the maxima
    }
    max_of_min = mx3(mx3(ra[0], ra[1], ra[2]), mx3(ra[3], ra[4], ra[5]), __vmaxs2(ra[6], ra[7]));
    min_of_max = mn3(mn3(rb[0], rb[1], rb[2]), mn3(rb[3], rb[4], rb[5]), __vmins2(rb[6], rb[7]));
}

// scores of the 4 pixels x0..x0+3 of the centre row r[3] (window rows r[0..6] = y-3..y+3), packed u8x4
template <bool HF, int FA> __device__ __forceinline__ uint32_t fast_score4(const Row6 (&r)[7], int th) {
    // circle (dx,dy), OpenCV order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
    const uint32_t e2_5 = E2(r[5]), em1_5 = Em1(r[5]), e2_1 = E2(r[1]), em1_1 = Em1(r[1]);
    const uint32_t em1_6 = Em1(r[6]), e2_6 = E2(r[6]), em1_0 = Em1(r[0]), e2_0 = E2(r[0]);
    const uint32_t e3_4 = E3(r[4]), e3_3 = E3(r[3]), e3_2 = E3(r[2]);
    const uint32_t em2_4 = Em2(r[4]), em2_3 = Em2(r[3]), em2_2 = Em2(r[2]);
    uint32_t P[16], Q[16];
    // pixels (x0, x0+2)
    P[0] = E0(r[6]);  P[1] = E1(r[6]);  P[2] = e2_5;       P[3] = e3_4;
    P[4] = e3_3;      P[5] = e3_2;      P[6] = e2_1;       P[7] = E1(r[0]);
    P[8] = E0(r[0]);  P[9] = em1_0;     P[10] = Em2(r[1]); P[11] = Em3(r[2]);
    P[12] = Em3(r[3]); P[13] = Em3(r[4]); P[14] = Em2(r[5]); P[15] = em1_6;
    // pixels (x0+1, x0+3): same offsets shifted by one column
    Q[0] = E1(r[6]);  Q[1] = e2_6;      Q[2] = E3(r[5]);   Q[3] = E4(r[4]);
    Q[4] = E4(r[3]);  Q[5] = E4(r[2]);  Q[6] = E3(r[1]);   Q[7] = e2_0;
    Q[8] = E1(r[0]);  Q[9] = E0(r[0]);  Q[10] = em1_1;     Q[11] = em2_2;
    Q[12] = em2_3;    Q[13] = em2_4;    Q[14] = em1_5;     Q[15] = E0(r[6]);
    uint32_t amaxP, bminP, amaxQ, bminQ;
    arc_extrema<HF, FA>(P, amaxP, bminP);
    arc_extrema<HF, FA>(Q, amaxQ, bminQ);
    const uint32_t cP = E0(r[3]), cQ = E1(r[3]);
    // best = max( c - min_arcs(max_arc), max_arcs(min_arc) - c )   (per 16-bit lane, signed)
    const uint32_t bestP = __vmaxs2(__vsub2(cP, bminP), __vsub2(amaxP, cP));
    const uint32_t bestQ = __vmaxs2(__vsub2(cQ, bminQ), __vsub2(amaxQ, cQ));
    // score = best - 1 where best > th, else 0 -- still on 16-bit pairs; then interleave (px0, px2) and (px1, px3) into bytes
    const uint32_t th2 = (uint32_t)th * 0x00010001u;
    const uint32_t sP = __vsub2(bestP, 0x00010001u) & __vcmpgts2(bestP, th2);
    const uint32_t sQ = __vsub2(bestQ, 0x00010001u) & __vcmpgts2(bestQ, th2);
    return __byte_perm(sP, sQ, 0x6240);
}

template <bool HF, int FA, int ROWS> __global__ void __launch_bounds__(128, 5) fast_score_kernel(DevPtrs d, FrameLayout fl, LevelBlocks lb, int th) {
    if (blockIdx.x == 0) {
        // this frame's survivor counters and cell flags, consumed by fast_nms_kernel two launches later on the same
        // stream (saves two memset nodes per batch)
        int32_t* cc = d.cell_count + (size_t)blockIdx.y * fl.n_cells;
        for (int i = threadIdx.x; i < fl.n_cells; i += blockDim.x) cc[i] = 0;
        if (threadIdx.x < kMaxLevels) d.level_raw[(size_t)blockIdx.y * kMaxLevels + threadIdx.x] = 0;
    }
    int first_block;
    const int level = level_of_block(lb, (int)blockIdx.x, first_block);
    const LevelGeom g = d.geom[level];
    // strips: x0 = 16 + 4*sx covers [16, W-19); y0 = 19 + kFastRows*sy covers [19, H-19)
    const int nsx = (g.w - kEdge - kMinBorder + 3) >> 2;
    const int id = (blockIdx.x - first_block) * blockDim.x + threadIdx.x;
    const int sy = id / nsx;
    const int x0 = kMinBorder + 4 * (id - sy * nsx);
    const int y0 = kEdge + sy * ROWS;
    if (y0 >= g.h - kEdge) return;
    int pitch;
    const uint8_t* src = level_roi(d, fl, g, level, blockIdx.y, pitch) + (x0 - 4);   // x0 >= 16: never leaves the row
    uint8_t* dst = d.score + (size_t)blockIdx.y * fl.splane_bytes + g.splane_off + x0;
    // byte mask of the columns inside [19, W-19)
    uint32_t colmask = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (x0 + b >= kEdge && x0 + b < g.w - kEdge) colmask |= 0xffu << (8 * b);

    // The 7-row window lives in registers and turns by one row per output row.  The row loop is rolled in blocks of seven
    // fully unrolled rows: after seven rows the window is back in the registers it started in (no copies), and the body is
    // half the code of a 14-row unroll -- the straight-line version streamed 56 KB of instructions once per warp and a
    // fifth of its stall samples were instruction-cache misses.  The three words of the NEXT row are requested a whole row
    // of arithmetic ahead of their use (the load of a row used to sit right in front of its first use: long-scoreboard
    // stalls were the top reason).
    Row6 r[7];
#pragma unroll
    for (int k = 0; k < 6; k++) r[k] = load_row6<HF>(src + (ptrdiff_t)(y0 - 3 + k) * pitch);
    uint32_t n0, n1, n2;
    {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(src + (ptrdiff_t)min(y0 + 3, g.h - 1) * pitch);
        n0 = w[0]; n1 = w[1]; n2 = w[2];
    }
    static_assert(ROWS % 7 == 0, "blocks of seven rows");
#pragma unroll 1
    for (int blk = 0; blk < ROWS / 7; blk++) {
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const int y = y0 + blk * 7 + j;
            r[6] = unpack_row6<HF>(n0, n1, n2);
            {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(src + (ptrdiff_t)min(y + 4, g.h - 1) * pitch);
                n0 = w[0]; n1 = w[1]; n2 = w[2];
            }
            const uint32_t s4 = fast_score4<HF, FA>(r, th) & colmask;
            if (y < g.h - kEdge) *reinterpret_cast<uint32_t*>(dst + (size_t)y * g.spitch) = s4;
#pragma unroll
            for (int k = 0; k < 6; k++) r[k] = r[k + 1];
        }
    }
}

int launch_fast_score(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, int min_th,
                      cudaStream_t s) {
    // rows per strip: 14 for batches, 7 for a few frames (twice as many threads of half the length: a single frame is a latency
    // problem, 114 CTAs of 14-row strips do not even fill the SMs)
    const int rows = n_frames >= 4 ? kFastRows : 7;
    LevelBlocks lb;
    int total = 0;
    const int threads = 128;
    for (int l = 0; l < fl.nlevels; l++) {
        lb.start[l] = total;
        const int nsx = (hg[l].w - kEdge - kMinBorder + 3) / 4;
        const int nsy = (hg[l].h - 2 * kEdge + rows - 1) / rows;
        total += (nsx * nsy + threads - 1) / threads;
    }
    for (int l = fl.nlevels; l <= kMaxLevels; l++) lb.start[l] = total;
    // ORBCUDA_FAST_FMA = number of the 24 two-input operations per side issued on the FMA pipe (0 keeps every min/max on the ALU
    // pipe; A/B switch, results are identical)
    static const int fa = [] { const char* e = getenv("ORBCUDA_FAST_FMA"); return e ? atoi(e) : kFastFmaDefault; }();
    const dim3 grid(total, n_frames);
    if (rows == 7) { fast_score_kernel<true, 8, 7><<<grid, threads, 0, s>>>(d, fl, lb, min_th); return 1; }
    switch (fa) {
        case 0: fast_score_kernel<false, 0, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
        case 12: fast_score_kernel<true, 12, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
        case 16: fast_score_kernel<true, 16, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
        case 20: fast_score_kernel<true, 20, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
        case 24: fast_score_kernel<true, 24, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
        default: fast_score_kernel<true, 8, kFastRows><<<grid, threads, 0, s>>>(d, fl, lb, min_th); break;
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------
// K3b: dense cell-aware non-max suppression + candidate emission.  A pixel survives if its score is
// strictly greater than its 8 neighbours, counting neighbours outside the pixel's own FAST cell interior
// as 0 (cv::FAST runs on the cell sub-image, so its NMS never sees them).  Cell interiors tile
// [19,W-19)x[19,H-19) with period (wCell,hCell) (R21 :773-807), so "is this pixel on its cell's
// left/right/top/bottom edge" is a function of (x-19) mod wCell and (y-19) mod hCell.  4 pixels (packed
// bytes) per thread, rows roll through registers.  Survivors (1-2 % of the pixels) are buffered in
// registers and appended to the level's list with ONE atomic per warp (shuffle prefix); a survivor
// reaching iniThFAST also raises its cell's flag (plain store).  The list order is arbitrary: the
// quadtree kernel applies the per-cell iniTh -> minTh fallback from the flags and resolves ties with an
// explicit (cell, raster) order key, so nothing downstream depends on it.
// ---------------------------------------------------------------------------------------------
constexpr int kNmsRows = 16;   // < 30 <= h_cell (FAST cells are at least 30 rows high, extractor.cu)

__global__ void __launch_bounds__(128) fast_nms_kernel(const uint8_t* __restrict__ score, int64_t plane_frame_bytes,
                                                       uint32_t* __restrict__ cand, int64_t cand_frame_entries,
                                                       int32_t* __restrict__ cell_flag, int n_cells,
                                                       int32_t* __restrict__ level_raw, const LevelGeom* __restrict__ geom,
                                                       int nlevels, LevelBlocks lb, int ini_th) {
    int first_block;
    const int level = level_of_block(lb, (int)blockIdx.x, first_block);
    const LevelGeom g = geom[level];
    const int lane = threadIdx.x & 31;
    const int nsx = (g.w - kEdge - kMinBorder + 3) >> 2;     // same strips as the score kernel: x0 = 16 + 4*sx
    const unsigned id = (blockIdx.x - first_block) * blockDim.x + threadIdx.x;
    const int sy = (int)(id / (unsigned)nsx);
    const int x0 = kMinBorder + 4 * ((int)id - sy * nsx);
    const int y0 = kEdge + sy * kNmsRows;
    const bool active = y0 < g.h - kEdge;
    uint32_t* list = cand + (size_t)blockIdx.y * cand_frame_entries + g.cand_off;
    int32_t* counter = level_raw + (size_t)blockIdx.y * kMaxLevels + level;
    const uint8_t* src = score + (size_t)blockIdx.y * plane_frame_bytes + g.splane_off + x0;
    // All 18 rows of the strip are requested before anything is computed: 54 independent 32-bit loads in flight per
    // thread.  The block barrier pins them there (without it ptxas sinks each row's loads next to their use to save
    // registers and every row pays the full memory latency: the kernel was latency-bound, ncu long-scoreboard stalls).
    // Strips past the last row clamp to row H-19 (exists in the plane) and are discarded.
    const int ylast = g.h - kEdge;   // row H-19 is never a centre row
    uint32_t W0[kNmsRows + 2], W1[kNmsRows + 2], W2[kNmsRows + 2];
#pragma unroll
    for (int r = 0; r < kNmsRows + 2; r++) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (ptrdiff_t)min(y0 - 1 + r, ylast) * g.spitch);
        W0[r] = __ldg(p - 1); W1[r] = __ldg(p); W2[r] = __ldg(p + 1);
    }
    __syncthreads();
    // survivor bits of the strip, 4 rows per word: pixel (x0+b, y0+4g+q) is bit 8 + 2q + (b&1) + 16*(b>>1) of km[g]
    uint32_t km[kNmsRows / 4] = {0, 0, 0, 0};
    // cell column of pixel x0+b = col0 + (b >= wrapb); cell row of strip row r = cy0 + (r > rb)
    const unsigned ux = (unsigned)(x0 - kEdge + g.w_cell);                 // >= w_cell - 3 > 0
    const unsigned qx = ux / (unsigned)g.w_cell;
    const int m0 = (int)(ux - qx * (unsigned)g.w_cell);                     // (x0 - 19) mod w_cell
    const int col0 = (int)qx - 1, wrapb = g.w_cell - m0;
    const unsigned uy = (unsigned)(y0 - kEdge);
    const int cy0 = (int)(uy / (unsigned)g.h_cell);
    const int ymod0 = (int)uy - cy0 * g.h_cell;
    // h_cell >= 30 > kNmsRows: a strip crosses at most one cell boundary.  rb = strip row that is the last row of
    // its cell; the row after it is the first row of the next cell.
    const int rb = g.h_cell - 1 - ymod0;
    const int rlast = g.h - kEdge - 1 - y0;            // last row of [19, H-19)
    if (active) {
        // Scores are handled as 16-bit pairs (the byte-wise SIMD intrinsics are emulated on sm_100a, the 16x2
        // min/max is one instruction): "even" words hold pixels (x0, x0+2), "odd" words (x0+1, x0+3).
        // Per-lane masks (0x00ff or 0): left / right neighbour in the same cell; KC: bit 8 of the lane if the pixel is in range.
        uint32_t MLe = 0, MLo = 0, MRe = 0, MRo = 0, KCe = 0, KCo = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int x = x0 + b;
            if (x >= kEdge && x < g.w - kEdge) {
                const int m = m0 + b - (b >= wrapb ? g.w_cell : 0);
                const uint32_t lane16 = 0xffu << (16 * (b >> 1));
                const bool l_ok = m != 0, r_ok = m != g.w_cell - 1 && x != g.w - kEdge - 1;
                if (b & 1) { KCo |= 0x100u << (16 * (b >> 1)); if (l_ok) MLo |= lane16; if (r_ok) MRo |= lane16; }
                else       { KCe |= 0x100u << (16 * (b >> 1)); if (l_ok) MLe |= lane16; if (r_ok) MRe |= lane16; }
            }
        }
        // per row: centre pairs, max(left, right) and max(left, centre, right) with out-of-cell horizontal neighbours zeroed
        uint32_t Ce[kNmsRows + 2], Co[kNmsRows + 2], LRe[kNmsRows + 2], LRo[kNmsRows + 2], He[kNmsRows + 2], Ho[kNmsRows + 2];
#pragma unroll
        for (int r = 0; r < kNmsRows + 2; r++) {
            Ce[r] = W1[r] & 0x00ff00ffu;                                     // (p0, p2)
            Co[r] = __byte_perm(W1[r], 0u, 0x4341);                          // (p1, p3)
            const uint32_t Le = __byte_perm(W0[r], W1[r], 0x0503) & MLe;     // (p-1, p1); the mask also clears bytes 1 and 3
            const uint32_t Ro = __byte_perm(W1[r], W2[r], 0x0402) & MRo;     // (p2, p4)
            const uint32_t Re = Co[r] & MRe, Lo = Ce[r] & MLo;
            LRe[r] = __vmaxs2(Le, Re); He[r] = __vimax3_s16x2(Le, Re, Ce[r]);
            LRo[r] = __vmaxs2(Lo, Ro); Ho[r] = __vimax3_s16x2(Lo, Ro, Co[r]);
        }
#pragma unroll
        for (int r = 0; r < kNmsRows; r++) {
            const bool top = (r == rb + 1) || (r == 0 && ymod0 == 0);
            const bool bot = (r == rb) || (r == rlast);
            uint32_t me = LRe[r + 1], mo = LRo[r + 1];
            if (!top) { me = __vmaxs2(me, He[r]); mo = __vmaxs2(mo, Ho[r]); }          // row 18 is never used: r = 0 there is a cell top
            if (!bot) { me = __vmaxs2(me, He[r + 2]); mo = __vmaxs2(mo, Ho[r + 2]); }
            // strict: 256 + m - c keeps bit 8 of the lane set unless c > m (equal neighbours suppress each other)
            const uint32_t ke = ~(me + 0x01000100u - Ce[r + 1]) & KCe;
            const uint32_t ko = ~(mo + 0x01000100u - Co[r + 1]) & KCo;
            if (r <= rlast) km[r >> 2] += (ke + 2u * ko) << (2 * (r & 3));
        }
    }
    // warp-aggregated append: one atomic per warp reserves room for all its survivors
    const int mine = __popc(km[0]) + __popc(km[1]) + __popc(km[2]) + __popc(km[3]);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(counter, total);
    int pos = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
    if (mine == 0) return;
    int32_t* flags = cell_flag + (size_t)blockIdx.y * n_cells + g.cell_base + cy0 * g.n_cols + col0;
    const uint8_t* srow = src + (size_t)y0 * g.spitch;
    const uint32_t packed0 = (uint32_t)(x0 - kMinBorder) | ((uint32_t)(y0 - kMinBorder) << 12);
    // one loop over the strip's 64 survivor bits (16 per group of four rows, packed: bit t of a group = pixel b = (t & 1) +
    // 2 * (t >> 3), row (t & 7) >> 1): a warp runs as many iterations as its busiest lane has survivors -- four loops, one per
    // group, ran the sum of the four per-group maxima, about twice as many
    auto pack16 = [](uint32_t m) { return ((m >> 8) & 0xffu) | ((m >> 16) & 0xff00u); };
    unsigned long long m = (unsigned long long)(pack16(km[0]) | (pack16(km[1]) << 16)) |
                           ((unsigned long long)(pack16(km[2]) | (pack16(km[3]) << 16)) << 32);
    while (m) {
        const int T = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int r = (T >> 4) * 4 + ((T & 7) >> 1), b = (T & 1) + 2 * ((T >> 3) & 1);
        const uint32_t sc = srow[r * g.spitch + b];
        list[pos++] = (packed0 + (uint32_t)b + ((uint32_t)r << 12)) | (sc << 24);
        if ((int)sc >= ini_th) flags[(r > rb ? g.n_cols : 0) + (b >= wrapb ? 1 : 0)] = 1;
    }
}

int launch_fast_cells(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, int ini_th,
                      cudaStream_t s) {
    LevelBlocks lb;
    int total = 0;
    const int nthreads = 128;
    for (int l = 0; l < fl.nlevels; l++) {
        lb.start[l] = total;
        const int nsx = (hg[l].w - kEdge - kMinBorder + 3) / 4;
        const int nsy = (hg[l].h - 2 * kEdge + kNmsRows - 1) / kNmsRows;
        total += (nsx * nsy + nthreads - 1) / nthreads;
    }
    for (int l = fl.nlevels; l <= kMaxLevels; l++) lb.start[l] = total;
    // d.cell_count and d.level_raw were zeroed by fast_score_kernel (launched before this kernel on the same stream)
    fast_nms_kernel<<<dim3(total, n_frames), nthreads, 0, s>>>(d.score, fl.splane_bytes, d.cand, fl.cand_entries,
                                                               d.cell_count, fl.n_cells, d.level_raw, d.geom, fl.nlevels, lb,
                                                               ini_th);
    return 1;
}

}  // namespace orbcuda
