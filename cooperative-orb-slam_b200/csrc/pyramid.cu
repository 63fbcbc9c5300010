// pyramid.cu -- K1: scale pyramid with fused REFLECT_101 border (replaces ORBextractor::ComputePyramid,
// R21/src/ORBextractor.cc:1107-1132) and K2: 7x7 sigma=2 Gaussian blur (replaces the GaussianBlur call
// at R21 :1085-1086).  uint8 planes, 4 pixels (one 32-bit word) per thread, HBM/L2 bound.
//
// Layout: a padded plane stores level pixel (x,y) at byte (y+19)*pitch + 32 + x; the 19-pixel
// REFLECT_101 border of cv::copyMakeBorder occupies columns [13,32) and [32+w, 51+w).  The interior
// origin is therefore 32-byte aligned and every row is `pitch` (multiple of 64) bytes.
#include "internal.h"

namespace orbcuda {

__device__ __forceinline__ int reflect101(int p, int len) {
    // cv::borderInterpolate(BORDER_REFLECT_101); |overshoot| <= 22 < len for every supported level
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}

// ---------------------------------------------------------------------------------------------
// level 0: copyMakeBorder(image, temp, 19,19,19,19, BORDER_REFLECT_101)   (R21 :1126-1128)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_level0_kernel(const uint8_t* __restrict__ in, size_t in_frame_stride,
                                                         int in_pitch, uint8_t* __restrict__ pyr,
                                                         int64_t pyr_frame_bytes, LevelGeom g) {
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;   // word index; word 0 starts at plane column 12
    const int py = blockIdx.y * blockDim.y + threadIdx.y;   // padded row
    const int px0 = 12 + 4 * wx;
    if (px0 >= kXPad + g.w + kEdge || py >= g.plane_rows) return;
    const uint8_t* src = in + (size_t)blockIdx.z * in_frame_stride;
    uint8_t* dst = pyr + (size_t)blockIdx.z * pyr_frame_bytes + g.plane_off;
    const int iy = reflect101(py - kEdge, g.h);
    const uint8_t* srow = src + (size_t)iy * in_pitch;
    uint32_t v;
    const int ix0 = px0 - kXPad;
    if (ix0 >= 0 && ix0 + 3 < g.w && ((reinterpret_cast<uintptr_t>(srow + ix0) & 3) == 0)) {
        v = *reinterpret_cast<const uint32_t*>(srow + ix0);
    } else {
        v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) v |= (uint32_t)srow[reflect101(ix0 + b, g.w)] << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)py * g.pitch + px0) = v;
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
// live rows and the 4 results leave as one 32-bit store.  Pixels within 19 px of an edge are also stored to
// their REFLECT_101 mirror positions, so the border needs no second pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_resize_kernel(uint8_t* pyr, int64_t pyr_frame_bytes, LevelGeom gs,
                                                         LevelGeom gd, const ResizeTap* __restrict__ xtab,
                                                         const ResizeTap* __restrict__ ytab) {
    extern __shared__ __align__(16) unsigned char s_src[];   // [rs_rows][rs_cols]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kPyrTileW, y0 = blockIdx.y * kPyrTileH;
    const int tw = min(kPyrTileW, gd.w - x0), th = min(kPyrTileH, gd.h - y0);
    uint8_t* frame = pyr + (size_t)blockIdx.z * pyr_frame_bytes;
    const uint8_t* src = frame + gs.plane_off + (size_t)kEdge * gs.pitch + kXPad;   // ROI origin of level l-1
    uint8_t* dst = frame + gd.plane_off + (size_t)kEdge * gd.pitch + kXPad;          // ROI origin of level l
    const ResizeTap* xt = xtab + gd.xtab_off;
    const ResizeTap* yt = ytab + gd.ytab_off;
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
            const uint8_t* gp = src + (size_t)sy0 * gs.pitch + sxa + 4 * wcol;
            for (int r = tid >> 6; r < nrows; r += 4)
                reinterpret_cast<uint32_t*>(s_src + r * spitch)[wcol] =
                    *reinterpret_cast<const uint32_t*>(gp + (size_t)r * gs.pitch);
        }
    }
    const int X0 = x0 + 4 * lane;
    const int npx = max(0, min(4, tw - 4 * lane));
    int o0[4], o1[4], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeTap t = xt[min(X0 + i, gd.w - 1)];
        o0[i] = t.ofs - sxa; o1[i] = t.pad - sxa; c0[i] = t.c0; c1[i] = t.c1;
    }
    __syncthreads();
    if (npx == 0) return;
    auto hrow = [&](int r, int (&hh)[4]) {
        const uint8_t* S = s_src + r * spitch;
#pragma unroll
        for (int i = 0; i < 4; i++) hh[i] = (S[o0[i]] * c0[i] + S[o1[i]] * c1[i]) >> 4;
    };
    constexpr int kNone = -(1 << 20);
    const bool xedge = X0 <= kEdge || X0 + 3 >= gd.w - 1 - kEdge;
    int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    int r_lo = kNone, r_hi = kNone;
    const int dy_end = min(th, warp * 8 + 8);
    for (int dy = warp * 8; dy < dy_end; dy++) {
        const int Y = y0 + dy;
        const ResizeTap ty = yt[Y];
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
        const int b0 = ty.c0, b1 = ty.c1;
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = (uint32_t)((((b0 * lo[i]) >> 16) + ((b1 * hi[i]) >> 16) + 2) >> 2) & 0xffu;
        int ym = kNone;
        if (Y >= 1 && Y <= kEdge) ym = -Y;
        else if (Y >= gd.h - 1 - kEdge && Y <= gd.h - 2) ym = 2 * (gd.h - 1) - Y;
        uint8_t* row = dst + (ptrdiff_t)Y * gd.pitch;
        uint8_t* mrow = dst + (ptrdiff_t)(ym == kNone ? Y : ym) * gd.pitch;
        if (npx == 4) {
            const uint32_t pk = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
            *reinterpret_cast<uint32_t*>(row + X0) = pk;
            if (ym != kNone) *reinterpret_cast<uint32_t*>(mrow + X0) = pk;
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (i < npx) { row[X0 + i] = (uint8_t)v[i]; if (ym != kNone) mrow[X0 + i] = (uint8_t)v[i]; }
        }
        if (xedge) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int X = X0 + i;
                int xm = kNone;
                if (X >= 1 && X <= kEdge) xm = -X;
                else if (X >= gd.w - 1 - kEdge && X <= gd.w - 2) xm = 2 * (gd.w - 1) - X;
                if (i < npx && xm != kNone) {
                    row[xm] = (uint8_t)v[i];
                    if (ym != kNone) mrow[xm] = (uint8_t)v[i];
                }
            }
        }
    }
}

int launch_pyramid(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, size_t in_frame_stride,
                   cudaStream_t s) {
    int launches = 0;
    for (int l = 0; l < fl.nlevels; l++) {
        const LevelGeom& g = hg[l];
        if (l == 0) {
            const dim3 block(64, 4);
            const int nwords = (kXPad + g.w + kEdge - 12 + 3) / 4;
            const dim3 grid((nwords + block.x - 1) / block.x, (g.plane_rows + block.y - 1) / block.y, n_frames);
            pyr_level0_kernel<<<grid, block, 0, s>>>(d.in, in_frame_stride, fl.in_pitch, d.pyr, fl.pyr_bytes, g);
        } else {
            const size_t smem = (size_t)g.rs_rows * g.rs_cols;
            if (smem > 48 * 1024 || g.rs_cols > 256) return -1;   // scale factors this large are not supported
            const dim3 grid((g.w + kPyrTileW - 1) / kPyrTileW, (g.h + kPyrTileH - 1) / kPyrTileH, n_frames);
            pyr_resize_kernel<<<grid, 256, smem, s>>>(d.pyr, fl.pyr_bytes, hg[l - 1], g, d.xtab, d.ytab);
        }
        launches++;
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------
// K2 blur.  cv::GaussianBlur(7x7, sigma 2) on CV_8U uses an 8.8 fixed-point separable kernel
// [18,34,48,56,48,34,18] with one rounding: (sum + 32768) >> 16.  The source is clone() of the level
// ROI with BORDER_REFLECT_101, which is exactly what the padded plane already holds around the ROI,
// so the kernel is a plain 7x7 convolution over the padded plane.
//
// Each thread owns a 4-pixel-wide column strip and slides down kBlurRows rows keeping the last seven
// horizontal sums in registers.  Horizontal pass on 16-bit pairs: a 32-bit IMAD does two columns
// (row sums <= 255*256 fit 16 bits).  Vertical pass in 32 bit with the symmetric-tap factoring.
// ---------------------------------------------------------------------------------------------
constexpr int kBlurRows = 16;


__device__ __forceinline__ uint32_t funnel16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); }

// horizontal 7-tap sums of 4 adjacent pixels x0..x0+3 from the words covering x0-4 .. x0+7;
// returns (h(x0) | h(x0+2)<<16) in .x and (h(x0+1) | h(x0+3)<<16) in .y
__device__ __forceinline__ uint2 blur_hsum4(uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t A0 = w0 & 0x00ff00ffu, A1 = (w0 >> 8) & 0x00ff00ffu;
    const uint32_t B0 = w1 & 0x00ff00ffu, B1 = (w1 >> 8) & 0x00ff00ffu;
    const uint32_t C0 = w2 & 0x00ff00ffu, C1 = (w2 >> 8) & 0x00ff00ffu;
    // E[t] = (p(x0+t), p(x0+t+2)) for t=-3..4
    const uint32_t Em3 = A1, Em2 = funnel16(A0, B0), Em1 = funnel16(A1, B1), E0 = B0, E1 = B1,
                   E2 = funnel16(B0, C0), E3 = funnel16(B1, C1), E4 = C0;
    uint2 r;
    r.x = 18u * (Em3 + E3) + 34u * (Em2 + E2) + 48u * (Em1 + E1) + 56u * E0;
    r.y = 18u * (Em2 + E4) + 34u * (Em1 + E3) + 48u * (E0 + E2) + 56u * E1;
    return r;
}

__global__ void __launch_bounds__(128) blur7_kernel(const uint8_t* __restrict__ pyr, int64_t pyr_frame_bytes,
                                                    uint8_t* __restrict__ blur, int64_t blur_frame_bytes,
                                                    const LevelGeom* __restrict__ geom, int nlevels, LevelBlocks lb) {
    int level = 0;
    while (level + 1 < nlevels && (int)blockIdx.x >= lb.start[level + 1]) level++;
    const LevelGeom g = geom[level];
    // strips of a level are flattened so every block is full whatever the level width
    const int nsx = (g.w + 3) >> 2;
    const int id = (blockIdx.x - lb.start[level]) * blockDim.x + threadIdx.x;
    const int sy = id / nsx;
    const int x0 = 4 * (id - sy * nsx);
    const int y0 = sy * kBlurRows;
    if (y0 >= g.h) return;
    const uint8_t* src = pyr + (size_t)blockIdx.y * pyr_frame_bytes + g.plane_off + (size_t)kEdge * g.pitch + kXPad;
    uint8_t* dst = blur + (size_t)blockIdx.y * blur_frame_bytes + g.splane_off;

    // all 22 source rows of the strip are fetched up front (66 independent 32-bit loads in flight per thread),
    // then reduced to horizontal sums; the vertical pass slides over them
    uint32_t w0[kBlurRows + 6], w1[kBlurRows + 6], w2[kBlurRows + 6];
#pragma unroll
    for (int k = 0; k < kBlurRows + 6; k++) {
        const int y = min(y0 - 3 + k, g.h + 2);   // rows up to h+2 exist in the padded plane (border 19)
        const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (ptrdiff_t)y * g.pitch + x0 - 4);
        w0[k] = p[0]; w1[k] = p[1]; w2[k] = p[2];
    }
    uint2 hs[kBlurRows + 6];   // hs[k] = horizontal sums of row (y0-3+k)
#pragma unroll
    for (int k = 0; k < kBlurRows + 6; k++) hs[k] = blur_hsum4(w0[k], w1[k], w2[k]);
#pragma unroll
    for (int r = 0; r < kBlurRows; r++) {
        uint32_t o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            auto H = [&](int k) -> uint32_t {
                const uint32_t pr = (c & 1) ? hs[r + k].y : hs[r + k].x;
                return (c & 2) ? (pr >> 16) : (pr & 0xffffu);
            };
            const uint32_t s = 18u * (H(0) + H(6)) + 34u * (H(1) + H(5)) + 48u * (H(2) + H(4)) + 56u * H(3);
            o[c] = (s + 32768u) >> 16;
        }
        if (y0 + r < g.h)
            *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + r) * g.spitch + x0) =
                o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    }
}

int launch_blur(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, cudaStream_t s) {
    LevelBlocks lb;
    int total = 0;
    const int threads = 128;
    for (int l = 0; l < fl.nlevels; l++) {
        lb.start[l] = total;
        const int strips = ((hg[l].w + 3) / 4) * ((hg[l].h + kBlurRows - 1) / kBlurRows);
        total += (strips + threads - 1) / threads;
    }
    for (int l = fl.nlevels; l <= kMaxLevels; l++) lb.start[l] = total;
    blur7_kernel<<<dim3(total, n_frames), threads, 0, s>>>(d.pyr, fl.pyr_bytes, d.blur, fl.splane_bytes, d.geom,
                                                           fl.nlevels, lb);
    return 1;
}

}  // namespace orbcuda
