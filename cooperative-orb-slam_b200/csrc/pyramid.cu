// pyramid.cu -- K1: scale pyramid with fused REFLECT_101 border (replaces ORBextractor::ComputePyramid,
// R21/src/ORBextractor.cc:1107-1132) and K2: 7x7 sigma=2 Gaussian blur (replaces the GaussianBlur call
// at R21 :1085-1086).  uint8 planes, 4 pixels (one 32-bit word) per thread, HBM/L2 bound.
//
// Layout: level 0 is the input image itself; a level l >= 1 is stored in a padded plane with pixel (x,y) at
// byte (y+19)*pitch + 32 + x (interior origin 32-byte aligned, pitch a multiple of 64).  The 19-pixel
// REFLECT_101 border that cv::copyMakeBorder adds (R21 :1122-1128) is NOT materialised on the device: FAST,
// IC_Angle, the descriptor sampler and the stereo SAD never leave the level interior, the blur applies
// REFLECT_101 itself, and orbx_download_level rebuilds the border on the host for mvImagePyramid consumers.
#include "internal.h"

#include <algorithm>
#include <cstdlib>

namespace orbcuda {

__device__ __forceinline__ int reflect101(int p, int len) {
    // cv::borderInterpolate(BORDER_REFLECT_101); |overshoot| <= 22 < len for every supported level
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}

// ---------------------------------------------------------------------------------------------
// level l>0: resize(level l-1 ROI, INTER_LINEAR) + copyMakeBorder(REFLECT_101|ISOLATED)
// (R21 :1118-1123).  cv::resize 8-bit linear: 11-bit fixed-point taps, int32 horizontal pass,
// vertical pass (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  Taps are computed on the host
// exactly as OpenCV does (float -> saturate_cast<short>) and read from a table.
//
// One CTA produces a kPyrTileW x kPyrTileH tile of the level interior.  The source footprint is staged in
// shared memory with coalesced 32-bit loads.  Each thread owns 4 adjacent output columns (its 8 horizontal
// taps stay in registers) and each warp streams down 8 output rows: the horizontal pass of a source row
// ((r>>4), 16 bits) is computed once and reused by the next output row, the vertical pass combines the two
// live rows and the 4 results leave as one 32-bit store.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_resize_kernel(DevPtrs d, FrameLayout fl, int level) {
    extern __shared__ __align__(16) unsigned char s_src[];   // [rs_rows][rs_cols]
    const LevelGeom gs = d.geom[level - 1], gd = d.geom[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kPyrTileW, y0 = blockIdx.y * kPyrTileH;
    const int tw = min(kPyrTileW, gd.w - x0), th = min(kPyrTileH, gd.h - y0);
    int src_pitch;
    const uint8_t* src = level_roi(d, fl, gs, level - 1, blockIdx.z, src_pitch);    // pixel (0,0) of level l-1
    uint8_t* dst = d.pyr + (size_t)blockIdx.z * fl.pyr_bytes + gd.plane_off + (size_t)kEdge * gd.pitch + kXPad;
    const ResizeTap* xt = d.xtab + gd.xtab_off;
    const ResizeTap* yt = d.ytab + gd.ytab_off;
    // source footprint of the tile
    const int sxa = xt[x0].ofs & ~3;
    const int sxe = xt[x0 + tw - 1].pad;
    const int sy0 = yt[y0].ofs;
    const int nrows = yt[y0 + th - 1].pad - sy0 + 1;
    const int nwords = (sxe - sxa + 4) >> 2;       // <= 64 (checked on the host)
    const int spitch = gd.rs_cols;                 // bytes per staged row (multiple of 4)
    {
        const int wcol = tid & 63;
        if (wcol < nwords) {
            const uint8_t* gp = src + (size_t)sy0 * src_pitch + sxa + 4 * wcol;
            for (int r = tid >> 6; r < nrows; r += 4)
                reinterpret_cast<uint32_t*>(s_src + r * spitch)[wcol] =
                    *reinterpret_cast<const uint32_t*>(gp + (size_t)r * src_pitch);
        }
    }
    const int X0 = x0 + 4 * lane;
    const int npx = max(0, min(4, tw - 4 * lane));
    // horizontal taps of my 4 columns: byte offset of the left source pixel and both weights packed as 16-bit pairs.
    // The right source pixel is always the next byte: where OpenCV clamps it (pad == ofs at the right edge) the
    // fraction is zero, so c1 == 0 and whatever byte follows contributes nothing.
    int o0[4];
    uint32_t cw[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeTap t = xt[min(X0 + i, gd.w - 1)];
        o0[i] = t.ofs - sxa;
        cw[i] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
    }
    __syncthreads();
    if (npx == 0) return;
    auto hrow = [&](int r, int (&hh)[4]) {
        const uint8_t* S = s_src + r * spitch;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t two = (uint32_t)S[o0[i]] | ((uint32_t)S[o0[i] + 1] << 8);
            int acc;      // S[o0]*c0 + S[o0+1]*c1: 16-bit weights x unsigned bytes
            asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(acc) : "r"(cw[i]), "r"(two), "r"(0));
            hh[i] = acc >> 4;
        }
    };
    constexpr int kNone = -(1 << 20);
    int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    int r_lo = kNone, r_hi = kNone;
    const int dy_end = min(th, warp * 8 + 8);
    for (int dy = warp * 8; dy < dy_end; dy++) {
        const int Y = y0 + dy;
        const uint2 traw = *reinterpret_cast<const uint2*>(yt + Y);      // {ofs, c0, c1, pad} in one load
        ResizeTap ty;
        ty.ofs = (int16_t)(traw.x & 0xffffu); ty.c0 = (int16_t)(traw.x >> 16); ty.c1 = (int16_t)(traw.y & 0xffffu); ty.pad = (int16_t)(traw.y >> 16);
        const int r0 = ty.ofs - sy0, r1 = ty.pad - sy0;
        if (r0 != r_lo) {
            if (r0 == r_hi) {
#pragma unroll
                for (int i = 0; i < 4; i++) lo[i] = hi[i];
            } else {
                hrow(r0, lo);
            }
            r_lo = r0;
        }
        if (r1 != r_hi) {
            if (r1 == r_lo) {
#pragma unroll
                for (int i = 0; i < 4; i++) hi[i] = lo[i];
            } else {
                hrow(r1, hi);
            }
            r_hi = r1;
        }
        // ((b0*lo) >> 16) + ((b1*hi) >> 16) + 2 as two multiply-high-adds on weights pre-shifted by 16 (the products are < 2^42,
        // each term is floored on its own exactly as in OpenCV): 2 instructions on the FMA pipe instead of 2 multiplies + 2 shifts
        // + 2 adds, and the result is < 1024 so no mask is needed after the final shift
        const int b0 = (int)ty.c0 << 16, b1 = (int)ty.c1 << 16;
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = (uint32_t)(__mulhi(b1, hi[i]) + __mulhi(b0, lo[i]) + 2) >> 2;
        uint8_t* row = dst + (ptrdiff_t)Y * gd.pitch;
        if (npx == 4) {
            *reinterpret_cast<uint32_t*>(row + X0) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < npx) row[X0 + i] = (uint8_t)v[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1, two-phase form (ORBCUDA_PYR=2; NOT the default).  The first kernel above spends ~49 thread-instruction slots per
// output pixel, most of them on the half-rate integer ALU pipe that bounds the whole extraction step.  This one does the
// same arithmetic (bit-identical output, same tests) in two passes over shared memory:
//   phase H: every source row of the tile's footprint is reduced ONCE to its horizontal sums (S[o]*c0 + S[o+1]*c1) >> 4,
//            two IMADs (FMA pipe) and one shift per tap, stored as int32 (one 16-byte store per four columns);
//   phase V: every output pixel = (((b0*lo) >> 16) + ((b1*hi) >> 16) + 2) >> 2 with the two products taken as
//            mul.hi.u32(b << 16, sum) (no shift), four pixels packed by multiply-adds and stored as one word.
// A thread owns one 4-column group for the whole tile (its eight tap weights and four source offsets stay in registers,
// every address advances by a constant) and the CTA's threads are laid out [RY][CGX] over (row, column group) with CGX =
// the tile's number of column groups, whatever that is -- warps simply wrap around rows, so level widths that are no
// multiple of 128 px cost nothing.  Tiles span the level width (split above ~640 px) and are as tall as 40 KB of sums
// allow (extractor.cu: t2_*).  The source footprint is staged with 16-byte loads.
// Measured on the B200 (round 2, ncu + CUDA events, 64-128 frames per launch): the inner loops are as lean as planned
// (27 instructions per four horizontal sums, ~40 per four output pixels) but a tile is only ~13 rows tall at 40 KB of sums,
// so the per-CTA set-up and the three barrier-separated phases dominate: 140 us per 64 frames for the seven levels
// against 128 us for the first kernel; walking six row tiles per CTA (kPyr2Band) removes the set-up but leaves too few
// CTAs on the small levels (0.26 ms per 128 frames on one stream against 0.196).  Kept as the measured alternative; a
// version that wins needs rolling row buffers and asynchronous staging inside long-running CTAs.
// ---------------------------------------------------------------------------------------------
constexpr int kPyr2Threads = 512;

constexpr int kPyr2Band = 6;       // row tiles one CTA walks (amortises the per-CTA set-up: divisions, tap staging, pointer bases)

__global__ void __launch_bounds__(kPyr2Threads) pyr_resize2_kernel(DevPtrs d, FrameLayout fl, int level) {
    extern __shared__ __align__(16) unsigned char smem[];
    const LevelGeom gs = d.geom[level - 1], gd = d.geom[level];
    const int tid = threadIdx.x, frame = blockIdx.y;
    const int tx = blockIdx.x % gd.t2_nx, band = blockIdx.x / gd.t2_nx;
    const int x0 = tx * gd.t2_w;
    const int tw = min(gd.t2_w, gd.w - x0);
    int src_pitch;
    const uint8_t* src = level_roi(d, fl, gs, level - 1, frame, src_pitch);
    uint8_t* dst = d.pyr + (size_t)frame * fl.pyr_bytes + gd.plane_off + (size_t)kEdge * gd.pitch + kXPad;
    const ResizeTap* xt = d.xtab + gd.xtab_off + x0;
    const int sxa = xt[0].ofs & ~15;                       // 16-byte aligned start of the staged columns
    const int nvec = (xt[tw - 1].pad - sxa + 16) >> 4;     // 16-byte vectors per staged row
    const int sp = gd.t2_cols;                             // bytes per staged row (multiple of 16, slack behind the last pixel)
    const int CGX = gd.t2_w >> 2;                          // column groups of a full tile = thread layout [RY][CGX]
    const int W4 = (tw + 3) >> 2;                          // column groups of this tile
    // shared memory: [horizontal sums int4 x t2_rows x CGX][y taps int4 x t2_h][x offsets int x t2_w][x weights u32 x t2_w][source]
    int4* s_h = reinterpret_cast<int4*>(smem);
    int4* s_y = s_h + gd.t2_rows * CGX;
    int* s_xo = reinterpret_cast<int*>(s_y + gd.t2_h);
    uint32_t* s_xc = reinterpret_cast<uint32_t*>(s_xo + gd.t2_w);
    unsigned char* s_src = reinterpret_cast<unsigned char*>(s_xc + gd.t2_w);
    // ---- per CTA: the horizontal taps of my column group, the staging role of this thread
    for (int i = tid; i < 4 * W4; i += kPyr2Threads) {
        const uint2 raw = *reinterpret_cast<const uint2*>(xt + min(i, tw - 1));       // {ofs, c0 | c1, pad}
        s_xo[i] = (int)(raw.x & 0xffffu) - sxa;
        s_xc[i] = (raw.x >> 16) | (raw.y << 16);                                       // c0 | c1 << 16
    }
    const int RV = kPyr2Threads / nvec;                    // nvec <= 64 (host: tiles <= ~1000 source bytes wide)
    const int vr = tid / nvec, vc = tid - vr * nvec;
    const int RY = kPyr2Threads / CGX;                     // >= 2 (host: CGX <= 176)
    const int ry = tid / CGX, cg = tid - ry * CGX;
    const bool mine = ry < RY && cg < W4;
    __syncthreads();
    int4 o = make_int4(0, 0, 0, 0);
    uint32_t c0x = 0, c1x = 0, c0y = 0, c1y = 0, c0z = 0, c1z = 0, c0w = 0, c1w = 0;
    if (mine) {
        o = reinterpret_cast<const int4*>(s_xo)[cg];
        const uint4 c = reinterpret_cast<const uint4*>(s_xc)[cg];
        c0x = c.x & 0xffffu; c1x = c.x >> 16; c0y = c.y & 0xffffu; c1y = c.y >> 16;
        c0z = c.z & 0xffffu; c1z = c.z >> 16; c0w = c.w & 0xffffu; c1w = c.w >> 16;
    }
    const int pstep = RY * sp, hstep = RY * CGX;
    const ptrdiff_t rstep = (ptrdiff_t)RY * gd.pitch;
    const size_t gstep = (size_t)RV * src_pitch;
    const int sstep = RV * sp;
    const int n_last = tw - 4 * cg;                        // >= 4: my four columns are a full word

    const int ty_end = min(gd.t2_ny, (band + 1) * kPyr2Band);
    for (int ty = band * kPyr2Band; ty < ty_end; ty++) {
        const int y0 = ty * gd.t2_h;
        const int th = min(gd.t2_h, gd.h - y0);
        const ResizeTap* yt = d.ytab + gd.ytab_off + y0;
        const int sy0 = yt[0].ofs;
        const int nrows = yt[th - 1].pad - sy0 + 1;
        // ---- stage the source footprint (thread = one 16-byte column of the footprint, rows strided) and the vertical taps
        if (vr < RV) {
            const uint8_t* gp = src + (size_t)(sy0 + vr) * src_pitch + sxa + 16 * vc;
            unsigned char* sp_ = s_src + vr * sp + 16 * vc;
            for (int r = vr; r < nrows; r += RV) {
                *reinterpret_cast<uint4*>(sp_) = *reinterpret_cast<const uint4*>(gp);
                gp += gstep; sp_ += sstep;
            }
        }
        for (int i = tid; i < th; i += kPyr2Threads) {
            const uint2 raw = *reinterpret_cast<const uint2*>(yt + i);
            const int ofs = (int)(raw.x & 0xffffu), c0 = (int)(raw.x >> 16), c1 = (int)(raw.y & 0xffffu), pad = (int)(raw.y >> 16);
            s_y[i] = make_int4((ofs - sy0) * CGX * 16, (pad - sy0) * CGX * 16, c0 << 16, c1 << 16);   // byte offsets into s_h
        }
        __syncthreads();
        // ---- phase H: my column group, source rows ry, ry + RY, ...
        if (mine) {
            const unsigned char* p = s_src + ry * sp;
            const unsigned char *px = p + o.x, *py = p + o.y, *pz = p + o.z, *pw = p + o.w;
            int4* ph = s_h + ry * CGX + cg;
            for (int r = ry; r < nrows; r += RY) {
                int4 hh;
                hh.x = (int)(px[0] * c0x + px[1] * c1x) >> 4;
                hh.y = (int)(py[0] * c0y + py[1] * c1y) >> 4;
                hh.z = (int)(pz[0] * c0z + pz[1] * c1z) >> 4;
                hh.w = (int)(pw[0] * c0w + pw[1] * c1w) >> 4;
                *ph = hh;
                px += pstep; py += pstep; pz += pstep; pw += pstep; ph += hstep;
            }
        }
        __syncthreads();
        // ---- phase V: my column group, output rows ry, ry + RY, ...
        if (mine) {
            uint8_t* row = dst + (ptrdiff_t)(y0 + ry) * gd.pitch + x0 + 4 * cg;
            const int4* pt = s_y + ry;
            const unsigned char* hb = reinterpret_cast<const unsigned char*>(s_h + cg);
            for (int dy = ry; dy < th; dy += RY) {
                const int4 t = *pt;
                const int4 lo = *reinterpret_cast<const int4*>(hb + t.x), hi = *reinterpret_cast<const int4*>(hb + t.y);
                const uint32_t b0 = (uint32_t)t.z, b1 = (uint32_t)t.w;
                const uint32_t v0 = (__umulhi(b0, (uint32_t)lo.x) + __umulhi(b1, (uint32_t)hi.x) + 2u) >> 2;
                const uint32_t v1 = (__umulhi(b0, (uint32_t)lo.y) + __umulhi(b1, (uint32_t)hi.y) + 2u) >> 2;
                const uint32_t v2 = (__umulhi(b0, (uint32_t)lo.z) + __umulhi(b1, (uint32_t)hi.z) + 2u) >> 2;
                const uint32_t v3 = (__umulhi(b0, (uint32_t)lo.w) + __umulhi(b1, (uint32_t)hi.w) + 2u) >> 2;
                if (n_last >= 4) {
                    *reinterpret_cast<uint32_t*>(row) = v0 + v1 * 256u + v2 * 65536u + v3 * 16777216u;     // every v <= 255
                } else {
                    row[0] = (uint8_t)v0;
                    if (n_last > 1) row[1] = (uint8_t)v1;
                    if (n_last > 2) row[2] = (uint8_t)v2;
                }
                row += rstep; pt += RY;
            }
        }
        __syncthreads();      // the next row tile overwrites the staged source and the taps
    }
}

// ---------------------------------------------------------------------------------------------
// K1, pair-staged form (ORBCUDA_PYR=3; NOT the default).  Same arithmetic as pyr_resize_kernel, bit-identical output (the
// extractor tests pass with it selected).  What the first kernel spent its instructions on (ncu --page source, 95 instructions per four output pixels in the row loop and as many
// again per CTA in set-up for only 8 rows per warp):
//   * two byte loads + a merge per horizontal tap  -> the source footprint is staged as PAIRS (two bytes per source pixel: the
//     pixel and its right neighbour), so a tap is ONE 16-bit load that is already the dp2a operand;
//   * twelve register moves per row to swap the "low" / "high" source rows -> the row step exists twice, once per role of
//     the two sum registers (template parameter), and the loop alternates between them: no copies;
//   * the vertical taps loaded from global memory at the top of every row -> requested one row ahead;
//   * LevelGeom fetched from global memory by every thread -> the few values needed arrive as a kernel parameter;
//   * 8 rows per warp -> tiles up to 128 rows tall (t3_h, the level height split evenly): half the set-up per output row and
//     better row utilisation on the small levels (134 rows = 2 x 72 instead of 3 x 64).
// Measured on the B200 (round 2, ncu + bench.py): 50.7 M warp instructions per 64 frames for the seven levels against 60.6 M
// (-16 %), 131 us against 135 us under ncu, 0.2008 ms against 0.2006 ms per 128 frames on one stream and the same 170 k
// frames/s over four streams.  The instructions it removes are IMAD / LDS / MOV (FMA pipe, LSU); the integer-ALU-pipe share
// (shifts, PRMT, compares, adds -- the pipe that bounds the extraction step as a whole) is unchanged, its
// taller tiles stage longer before they compute (long-scoreboard stalls 35 %) and hold 50 KB of shared memory (4 CTAs per SM).
// Kept as the measured alternative; the first kernel stays the default.
// ---------------------------------------------------------------------------------------------
struct Pyr3Args {
    const uint8_t* src; size_t src_frame; int src_pitch;     // level l-1, pixel (0,0) of frame 0
    uint8_t* dst; size_t dst_frame; int dst_pitch;           // level l, pixel (0,0) of frame 0
    const ResizeTap* xt; const ResizeTap* yt;                // taps of this level
    int w, h;                                                // destination size
    int th;                                                  // tile height (a multiple of 8)
    int row_bytes;                                           // bytes per staged row of pairs
};

__global__ void __launch_bounds__(256) pyr_resize3_kernel(Pyr3Args a) {
    extern __shared__ __align__(16) unsigned char s_pair[];   // [rows][row_bytes]: pair x at byte 2 x
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kPyrTileW, y0 = blockIdx.y * a.th;
    const int tw = min(kPyrTileW, a.w - x0), th = min(a.th, a.h - y0);
    const uint8_t* src = a.src + (size_t)blockIdx.z * a.src_frame;
    uint8_t* dst = a.dst + (size_t)blockIdx.z * a.dst_frame;
    const ResizeTap* __restrict__ xt = a.xt;
    const ResizeTap* __restrict__ yt = a.yt;
    // source footprint of the tile
    const int sxa = xt[x0].ofs & ~3;
    const int sxe = xt[x0 + tw - 1].pad;
    const int sy0 = yt[y0].ofs;
    const int nrows = yt[y0 + th - 1].pad - sy0 + 1;
    const int nwords = (sxe - sxa + 4) >> 2;       // <= 64 (checked on the host)
    {
        // a warp stages 32 consecutive words of one source row per pass (rows tid >> 6, +4, ...); the byte that follows a word
        // comes from the next lane's word (shuffle), for lane 31 from one more load
        const int wcol = tid & 63;
        const bool mine = wcol < nwords;
        const bool next_in_row = mine && sxa + 4 * wcol + 4 < a.src_pitch;      // the word after the row's last one is never needed (zero weight)
        const uint8_t* gp = src + (size_t)(sy0 + (tid >> 6)) * a.src_pitch + sxa + 4 * wcol;
        unsigned char* sp = s_pair + (tid >> 6) * a.row_bytes + 8 * wcol;
        const size_t gstep = (size_t)4 * a.src_pitch;
        const int sstep = 4 * a.row_bytes;
        for (int r = tid >> 6; r < nrows; r += 4, gp += gstep, sp += sstep) {
            const uint32_t w = mine ? *reinterpret_cast<const uint32_t*>(gp) : 0u;
            uint32_t wn = __shfl_down_sync(0xffffffffu, w, 1);
            if (lane == 31) wn = next_in_row ? *reinterpret_cast<const uint32_t*>(gp + 4) : 0u;
            // pairs (p0,p1) (p1,p2) | (p2,p3) (p3,p4)
            if (mine) *reinterpret_cast<uint2*>(sp) = make_uint2(__byte_perm(w, wn, 0x2110), __byte_perm(w, wn, 0x4332));
        }
    }
    const int X0 = x0 + 4 * lane;
    const int npx = max(0, min(4, tw - 4 * lane));
    // horizontal taps of my 4 columns: shared-memory address of the pair and both weights packed as 16-bit pairs.  Where OpenCV
    // clamps the right tap (pad == ofs at the right edge) the fraction is zero, so c1 == 0 and the pair's second byte
    // contributes nothing.
    uint32_t pa[4], cw[4];
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_pair);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint2 traw = *reinterpret_cast<const uint2*>(xt + min(X0 + i, a.w - 1));      // {ofs, c0 | c1, pad}
        pa[i] = s_base + 2u * (uint32_t)((int)(int16_t)(traw.x & 0xffffu) - sxa);
        cw[i] = (traw.x >> 16) | (traw.y << 16);
    }
    __syncthreads();
    if (npx == 0) return;
    const int rw = a.th >> 3;
    int dy = warp * rw;
    const int dy_end = min(th, dy + rw);
    if (dy >= dy_end) return;

    int HA[4] = {0, 0, 0, 0}, HB[4] = {0, 0, 0, 0};
    constexpr int kNone = -(1 << 20);
    int r_lo = kNone, r_hi = kNone;
    uint8_t* out = dst + (size_t)(y0 + dy) * a.dst_pitch + X0;
    uint2 tnext = *reinterpret_cast<const uint2*>(yt + y0 + dy);
    auto hrow = [&](int r, int (&hh)[4]) {
        const uint32_t ro = (uint32_t)(r * a.row_bytes);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t two;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(two) : "r"(pa[i] + ro));
            int acc;      // S[o]*c0 + S[o+1]*c1: 16-bit weights x unsigned bytes
            asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(acc) : "r"(cw[i]), "r"(two), "r"(0));
            hh[i] = acc >> 4;
        }
    };
    // ((b0*lo) >> 16) + ((b1*hi) >> 16) + 2 as multiply-high-adds on weights pre-shifted by 16 (each term floored on its own exactly
    // as in OpenCV); the result is < 1024, so no mask is needed after the final shift
    auto emit = [&](const int (&L)[4], const int (&H)[4], int b0, int b1) {
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = (uint32_t)(__mulhi(b1, H[i]) + __mulhi(b0, L[i]) + 2) >> 2;
        if (npx == 4) {
            *reinterpret_cast<uint32_t*>(out) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < npx) out[i] = (uint8_t)v[i];
        }
        out += a.dst_pitch;
        dy++;
    };
    // one output row with the low source row's sums in `lo` and the high row's in `hi`; returns true if the roles swapped.  Every
    // path names its registers statically (no run-time choice between the two arrays: that costs a dozen moves per row).
    auto row_step = [&](int (&lo)[4], int (&hi)[4]) -> bool {
        const uint2 traw = tnext;
        if (dy + 1 < dy_end) tnext = *reinterpret_cast<const uint2*>(yt + y0 + dy + 1);
        const int r0 = (int)(int16_t)(traw.x & 0xffffu) - sy0, r1 = (int)(int16_t)(traw.y >> 16) - sy0;
        const int b0 = (int)(traw.x & 0xffff0000u), b1 = (int)(traw.y << 16);      // weights << 16
        if (r0 == r_hi && r1 != r0) {
            // the usual step: the old high row becomes the low row, the new high row goes where the old low row was
            hrow(r1, lo);
            r_lo = r0; r_hi = r1;
            emit(hi, lo, b0, b1);
            return true;
        }
        if (r0 != r_lo) { hrow(r0, lo); r_hi = kNone; }
        if (r1 == r0) {
#pragma unroll
            for (int i = 0; i < 4; i++) hi[i] = lo[i];
        } else if (r1 != r_hi) hrow(r1, hi);
        r_lo = r0; r_hi = r1;
        emit(lo, hi, b0, b1);
        return false;
    };
    // state 0: low row in HA, high row in HB; state 1: the other way round
    bool state1 = false;
    while (dy < dy_end) {
        if (!state1) { if (row_step(HA, HB)) state1 = true; }
        else { if (row_step(HB, HA)) state1 = false; }
    }
}

// th_small > 0: tiles of that many rows instead of the level's t3_h (a few frames: many short CTAs instead of few long ones)
static int launch_pyramid_v3(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, cudaStream_t s, int th_small = 0) {
    static DeviceOnce once;
    if (!once.run([&] { return cudaFuncSetAttribute(pyr_resize3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) == cudaSuccess; }))
        return -1;
    int launches = 0;
    for (int l = 1; l < fl.nlevels; l++) {   // level 0 is the input image itself
        const LevelGeom& g = hg[l];
        const LevelGeom& gs = hg[l - 1];
        Pyr3Args a;
        if (l == 1) { a.src = d.in; a.src_frame = d.in_frame_stride; a.src_pitch = fl.in_pitch; }
        else { a.src = d.pyr + gs.plane_off + (size_t)kEdge * gs.pitch + kXPad; a.src_frame = (size_t)fl.pyr_bytes; a.src_pitch = gs.pitch; }
        a.dst = d.pyr + g.plane_off + (size_t)kEdge * g.pitch + kXPad; a.dst_frame = (size_t)fl.pyr_bytes; a.dst_pitch = g.pitch;
        a.xt = d.xtab + g.xtab_off; a.yt = d.ytab + g.ytab_off;
        a.w = g.w; a.h = g.h; a.row_bytes = g.t3_row_bytes;
        size_t smem = (size_t)g.t3_smem + 16;
        a.th = g.t3_h;
        if (th_small > 0 && th_small < g.t3_h) {
            // rows a tile of th_small output rows can touch: its taps advance by src/dst rows per output row (+ the two ends)
            a.th = th_small;
            const int rows = (int)((double)th_small * gs.h / g.h) + 4;
            smem = (size_t)std::min(rows, g.t3_rows) * g.t3_row_bytes + 16;
        }
        const dim3 grid((g.w + kPyrTileW - 1) / kPyrTileW, (g.h + a.th - 1) / a.th, n_frames);
        pyr_resize3_kernel<<<grid, 256, smem, s>>>(a);
        launches++;
    }
    return launches;
}

static int launch_pyramid_v1(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, cudaStream_t s) {
    int launches = 0;
    for (int l = 1; l < fl.nlevels; l++) {   // level 0 is the input image itself
        const LevelGeom& g = hg[l];
        const size_t smem = (size_t)g.rs_rows * g.rs_cols + 16;     // + slack: the (zero-weight) byte after the last staged pixel is read
        if (smem > 48 * 1024 || g.rs_cols > 256) return -1;   // scale factors this large are not supported
        const dim3 grid((g.w + kPyrTileW - 1) / kPyrTileW, (g.h + kPyrTileH - 1) / kPyrTileH, n_frames);
        pyr_resize_kernel<<<grid, 256, smem, s>>>(d, fl, l);
        launches++;
    }
    return launches;
}

int launch_pyramid(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, cudaStream_t s) {
    // ORBCUDA_PYR: 0 (default) = the first kernel for batches and the pair-staged kernel with 16-row tiles for a few frames (a single
    // frame is a latency problem: seven dependent launches of ~9 us each with 64-row tiles, 8 rows per warp one after the other);
    // 1 the first kernel, 2 the two-phase kernel, 3 the pair-staged kernel with tall tiles (A/B switches; results are identical)
    static const int variant = [] { const char* e = getenv("ORBCUDA_PYR"); return e ? atoi(e) : 0; }();
    bool v3_ok = true;
    for (int l = 1; l < fl.nlevels; l++) v3_ok = v3_ok && hg[l].t3_smem > 0;
    static const int th_small = [] { const char* e = getenv("ORBCUDA_PYR_TH"); return e ? atoi(e) : 16; }();
    if (variant == 0) return (n_frames < 4 && v3_ok) ? launch_pyramid_v3(d, fl, hg, n_frames, s, th_small) : launch_pyramid_v1(d, fl, hg, n_frames, s);
    if (variant == 3) return v3_ok ? launch_pyramid_v3(d, fl, hg, n_frames, s) : launch_pyramid_v1(d, fl, hg, n_frames, s);
    if (variant == 1) return launch_pyramid_v1(d, fl, hg, n_frames, s);
    int max_smem = 0;
    for (int l = 1; l < fl.nlevels; l++) {
        max_smem = std::max(max_smem, hg[l].t2_smem);
        if (hg[l].t2_smem <= 0) return launch_pyramid_v1(d, fl, hg, n_frames, s);     // a shape the two-phase tiling does not cover
    }
    if (max_smem > 160 * 1024) return launch_pyramid_v1(d, fl, hg, n_frames, s);
    static DeviceOnce once;
    if (!once.run([&] { return cudaFuncSetAttribute(pyr_resize2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == cudaSuccess; }))
        return -1;
    int launches = 0;
    for (int l = 1; l < fl.nlevels; l++) {
        const LevelGeom& g = hg[l];
        pyr_resize2_kernel<<<dim3(g.t2_nx * ((g.t2_ny + kPyr2Band - 1) / kPyr2Band), n_frames), kPyr2Threads, g.t2_smem, s>>>(d, fl, l);
        launches++;
    }
    return launches;
}

// Input repack: rows of an arbitrarily strided / unaligned image block into the aligned layout the kernels read
// (pitch a multiple of 16).  One aligned 32-bit word per thread, bytes gathered through L1.
__global__ void __launch_bounds__(128) repack_kernel(const uint8_t* __restrict__ src, size_t src_row, size_t src_frame,
                                                     uint8_t* __restrict__ dst, int dst_pitch, size_t dst_frame, int width) {
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * wx >= width) return;
    const uint8_t* s = src + blockIdx.z * src_frame + blockIdx.y * src_row + 4 * wx;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (4 * wx + b < width) v |= (uint32_t)s[b] << (8 * b);
    *reinterpret_cast<uint32_t*>(dst + blockIdx.z * dst_frame + (size_t)blockIdx.y * dst_pitch + 4 * wx) = v;
}

int launch_repack(const uint8_t* src, size_t src_row, size_t src_frame, uint8_t* dst, int dst_pitch, size_t dst_frame, int width,
                  int height, int n_frames, cudaStream_t s) {
    const dim3 grid(((width + 3) / 4 + 127) / 128, height, n_frames);
    repack_kernel<<<grid, 128, 0, s>>>(src, src_row, src_frame, dst, dst_pitch, dst_frame, width);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// K2 blur.  cv::GaussianBlur(7x7, sigma 2) on CV_8U uses an 8.8 fixed-point separable kernel
// [18,34,48,56,48,34,18] with one rounding: (sum + 32768) >> 16.  The source is clone() of the level
// ROI with BORDER_REFLECT_101: rows are reflected by index, and the two edge strips of a row gather
// their 12-byte window with reflected column indices.
//
// Each thread owns a 4-pixel-wide column strip and slides down kBlurRows rows keeping the last seven
// horizontal sums in registers.  Horizontal pass on 16-bit pairs: a 32-bit IMAD does two columns
// (row sums <= 255*256 fit 16 bits).  Vertical pass in 32 bit with the symmetric-tap factoring.
// ---------------------------------------------------------------------------------------------
constexpr int kBlurRows = 16;



// horizontal 7-tap sums of 4 adjacent pixels x0..x0+3 from the words covering x0-4 .. x0+7 (w0 = bytes x0-4..x0-1,
// w1 = x0..x0+3, w2 = x0+4..x0+7): the window of pixel i is the 8 bytes starting i+1 bytes into (w0, w1), i.e. two
// funnel shifts, and the 7 taps are two dp4a with byte weights (18,34,48,56) (48,34,18,0).  Sums <= 255*256 fit 16 bits.
__device__ __forceinline__ uint4 blur_hsum4(uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t k0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), k1 = 48u | (34u << 8) | (18u << 16);
    uint4 r;
    r.x = __dp4a(__funnelshift_r(w1, w2, 8), k1, __dp4a(__funnelshift_r(w0, w1, 8), k0, 0u));
    r.y = __dp4a(__funnelshift_r(w1, w2, 16), k1, __dp4a(__funnelshift_r(w0, w1, 16), k0, 0u));
    r.z = __dp4a(__funnelshift_r(w1, w2, 24), k1, __dp4a(__funnelshift_r(w0, w1, 24), k0, 0u));
    r.w = __dp4a(w2, k1, __dp4a(w1, k0, 0u));
    return r;
}

// One strip: 4 columns x kBlurRows rows.  EDGE strips (window x0-4 .. x0+7 leaves the row) gather their
// bytes with reflected column indices; they are numbered after all interior strips of the level so that
// whole warps take one path or the other.
template <bool EDGE, int ROWS>
__device__ __forceinline__ void blur_strip(const uint8_t* __restrict__ src, int pitch, uint8_t* __restrict__ dst, int spitch,
                                           int w, int h, int x0, int y0) {
    // all 22 source rows of the strip are fetched up front (66 independent 32-bit loads in flight per thread),
    // then reduced to horizontal sums; the vertical pass slides over them
    uint32_t w0[ROWS + 6], w1[ROWS + 6], w2[ROWS + 6];
    if (!EDGE && y0 >= 3 && y0 + ROWS + 3 <= h) {
        // no row of the window is reflected (all but the first and last strip row of a level): one multiply-add per row
        // address instead of the two reflections + 64-bit address arithmetic below (a quarter of the kernel's instructions)
        const uint8_t* base = src + (size_t)(y0 - 3) * pitch + (x0 - 4);
#pragma unroll
        for (int k = 0; k < ROWS + 6; k++) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (ptrdiff_t)k * pitch);
            w0[k] = p[0]; w1[k] = p[1]; w2[k] = p[2];
        }
    } else
#pragma unroll
    for (int k = 0; k < ROWS + 6; k++) {
        int y = y0 - 3 + k;
        y = y < 0 ? -y : y;                               // BORDER_REFLECT_101
        y = y >= h ? max(2 * h - 2 - y, 0) : y;           // (rows past h+2 feed outputs that are never stored)
        const uint8_t* row = src + (size_t)y * pitch;
        if (!EDGE) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(row + x0 - 4);
            w0[k] = p[0]; w1[k] = p[1]; w2[k] = p[2];
        } else {
            // Edge strips: a word of the window that lies inside the row is loaded whole, only the others are put together
            // byte by byte with reflected column indices (12 byte loads per row made these few strips 40 % of the blur time).
            // Left edge (x0 == 0): the window's first word is pixels (4, 3, 2, 1) -- one PRMT of the two words that follow.
            const uint32_t* p = reinterpret_cast<const uint32_t*>(row + x0);
            uint32_t a, b, c;
            if (x0 + 3 < w) b = p[0];
            else {
                b = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { int xb = x0 + i; xb = xb >= w ? max(2 * w - 2 - xb, 0) : xb; b |= (uint32_t)row[xb] << (8 * i); }
            }
            if (x0 + 7 < w) c = p[1];
            else {
                c = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) { int xc = x0 + 4 + i; xc = xc >= w ? max(2 * w - 2 - xc, 0) : xc; c |= (uint32_t)row[xc] << (8 * i); }
            }
            if (x0 >= 4) a = p[-1];
            else a = __byte_perm(b, c, 0x1234);          // x0 == 0 (levels are far wider than 8 px: b and c are plain words here)
            w0[k] = a; w1[k] = b; w2[k] = c;
        }
    }
    // Vertical pass with dp2a.  the horizontal sums of rows (y0-3+k) come one pixel per register; the sums of rows 2j
    // and 2j+1 of ONE column are paired in a register (pc[c][j], one PRMT each), so the 7 taps of an output pixel are 4
    // two-way dot products with byte weights -- even output rows r = 2m: pairs m..m+3 . (18,34) (48,56) (48,34) (18,0);
    // odd rows r = 2m+1: pairs m..m+3 . (0,18) (34,48) (56,48) (34,18).  sum < 2^24, result = byte 2 of (sum + 32768),
    // picked and packed by PRMT.
    static_assert(ROWS % 2 == 0, "row pairs");
    uint32_t pc[4][ROWS / 2 + 3];
#pragma unroll
    for (int j = 0; j < ROWS / 2 + 3; j++) {
        const uint4 h0 = blur_hsum4(w0[2 * j], w1[2 * j], w2[2 * j]);
        const uint4 h1 = blur_hsum4(w0[2 * j + 1], w1[2 * j + 1], w2[2 * j + 1]);
        pc[0][j] = __byte_perm(h0.x, h1.x, 0x5410);
        pc[1][j] = __byte_perm(h0.y, h1.y, 0x5410);
        pc[2][j] = __byte_perm(h0.z, h1.z, 0x5410);
        pc[3][j] = __byte_perm(h0.w, h1.w, 0x5410);
    }
    uint8_t* const out = dst + (size_t)y0 * spitch + x0;
    const uint32_t we0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // even rows: pairs m, m+1
    const uint32_t we1 = 48u | (34u << 8) | (18u << 16) | (0u << 24);      //            pairs m+2, m+3
    const uint32_t wo0 = 0u | (18u << 8) | (34u << 16) | (48u << 24);      // odd rows:  pairs m, m+1
    const uint32_t wo1 = 56u | (48u << 8) | (34u << 16) | (18u << 24);     //            pairs m+2, m+3
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int m = r >> 1;
        const uint32_t wA = (r & 1) ? wo0 : we0, wB = (r & 1) ? wo1 : we1;
        uint32_t o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t sum;
            asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(pc[c][m]), "r"(wA), "r"(32768u));
            asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(pc[c][m + 1]), "r"(wA), "r"(sum));
            asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(pc[c][m + 2]), "r"(wB), "r"(sum));
            asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(sum) : "r"(pc[c][m + 3]), "r"(wB), "r"(sum));
            o[c] = sum;
        }
        if (y0 + r < h)
            *reinterpret_cast<uint32_t*>(out + (ptrdiff_t)r * spitch) =
                __byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410);
    }
}

// One launch: the interior strips (4 <= x0 <= w-8) of every level first, then the 2-3 edge strips per strip row of every
// level (blocks >= lb_edge.start[0] -- whole blocks take one path or the other).
template <int ROWS>
__global__ void __launch_bounds__(128) blur7_kernel(DevPtrs d, FrameLayout fl, LevelBlocks lb, LevelBlocks lb_edge) {
    const bool edge = (int)blockIdx.x >= lb_edge.start[0];
    int first_block;
    const int level = edge ? level_of_block(lb_edge, (int)blockIdx.x, first_block) : level_of_block(lb, (int)blockIdx.x, first_block);
    const LevelGeom g = d.geom[level];
    // strips of a level are flattened so every block is full whatever the level width
    const int nsx = (g.w + 3) >> 2;
    const int ni = max((g.w - 8) >> 2, 0);            // interior strips per row: x0 = 4, 8, ..., 4*ni
    const int per_row = edge ? nsx - ni : ni;
    const int nsy = (g.h + ROWS - 1) / ROWS;
    const int id = (blockIdx.x - first_block) * blockDim.x + threadIdx.x;
    if (id >= per_row * nsy) return;
    int pitch;
    const uint8_t* src = level_roi(d, fl, g, level, blockIdx.y, pitch);
    uint8_t* dst = d.blur + (size_t)blockIdx.y * fl.splane_bytes + g.splane_off;
    const int sy = id / per_row, k = id - sy * per_row;
    if (edge) blur_strip<true, ROWS>(src, pitch, dst, g.spitch, g.w, g.h, k == 0 ? 0 : 4 * (ni + k), sy * ROWS);
    else blur_strip<false, ROWS>(src, pitch, dst, g.spitch, g.w, g.h, 4 + 4 * k, sy * ROWS);
}

int launch_blur(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, cudaStream_t s) {
    // Rows per strip: 16 for batches (fewest redundant window rows: 22 loaded for 16 written); 4 for a few frames -- a single frame
    // has only ~15 k strips of 16 rows (one CTA per SM, each thread a serial chain of 16 rows: 18.7 us); four rows per strip are
    // four times as many threads of a quarter of the length (ORBCUDA_BLUR_ROWS forces one)
    static const int forced = [] { const char* e = getenv("ORBCUDA_BLUR_ROWS"); return e ? atoi(e) : 0; }();
    int rows = forced ? forced : (n_frames >= 4 ? kBlurRows : 4);
    if (rows != 4 && rows != 8) rows = kBlurRows;
    const int threads = 128;
    LevelBlocks lb[2];
    int total = 0;
    for (int edge = 0; edge < 2; edge++) {
        for (int l = 0; l < fl.nlevels; l++) {
            lb[edge].start[l] = total;
            const int nsx = (hg[l].w + 3) / 4, ni = std::max((hg[l].w - 8) / 4, 0);
            const int strips = (edge ? nsx - ni : ni) * ((hg[l].h + rows - 1) / rows);
            total += (strips + threads - 1) / threads;
        }
        for (int l = fl.nlevels; l <= kMaxLevels; l++) lb[edge].start[l] = total;
    }
    if (total == 0) return 0;
    if (rows == 4) blur7_kernel<4><<<dim3(total, n_frames), threads, 0, s>>>(d, fl, lb[0], lb[1]);
    else if (rows == 8) blur7_kernel<8><<<dim3(total, n_frames), threads, 0, s>>>(d, fl, lb[0], lb[1]);
    else blur7_kernel<kBlurRows><<<dim3(total, n_frames), threads, 0, s>>>(d, fl, lb[0], lb[1]);
    return 1;
}

}  // namespace orbcuda
