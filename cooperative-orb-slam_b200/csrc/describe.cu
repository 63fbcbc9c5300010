// describe.cu -- K5: IC_Angle orientation + steered rBRIEF descriptor, fused, one warp per key point
// (replaces computeOrientation/IC_Angle R21/src/ORBextractor.cc:472-479,:77-104 and
// computeDescriptors/computeOrbDescriptor :1034-1041,:108-147, plus the level -> image coordinate
// scaling and level-major concatenation of operator() :1072-1104).
//
// Float parity: fastAtan2 is OpenCV's degree-7 polynomial evaluated in strict fp32 with separate
// multiplies and adds (__fmul_rn/__fadd_rn: no FMA contraction); the pattern rotation is
// x*b + y*a / x*a - y*b again without FMA and rounded with cvRound == round-half-even
// (__float2int_rn).  cos/sin are evaluated in double and rounded once to float (the reference calls
// libm cosf/sinf, which is not correctly rounded and differs between glibc builds).
#include "internal.h"

#include <algorithm>

namespace orbcuda {

__constant__ int8_t c_pattern[1024] = {
#include "orb_pattern.inc"
};
// u_max of the 31x31 circular patch (R21 :452-469)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 scalar path (OpenCV core/mathfuncs_core atan_f32)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale);
    const float p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale);
    const float p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// round-half-even of |v| < 2^22 without the conversion pipe: adding 1.5*2^23 leaves rint(v) in the low mantissa bits
__device__ __forceinline__ int round_rn(float v) {
    return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4b400000;
}

constexpr int kDescWarps = 4;
constexpr int kPatchR = 18;                     // the pattern's largest radius is 18.38 (point (-13,-13)): |cvRound(rotated coordinate)| <= 18
constexpr int kPatchRows = 2 * kPatchR + 1;     // 37 rows of the blurred level around a key point
constexpr int kPatchStride = 64;                // bytes per staged row: four 16-byte chunks (the 128-bit stores of a quarter warp cover two whole rows: no bank conflicts)

__device__ __forceinline__ int dp4a_us(uint32_t px, uint32_t w, int acc) {   // 4 unsigned pixels x 4 signed weights
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(px), "r"(w), "r"(acc));
    return r;
}

// per key point of a warp's group: what the two cooperative phases need, broadcast through shared memory
struct __align__(16) KpWork {
    const uint8_t* img;     // un-blurred level, pixel (x - 15, y)                       } first 16 bytes: moments
    int pitch, cx;          // cx = x - (16-byte aligned start column of the staged rows)  }
    const uint8_t* blr;     // blurred level, row y - 18, the aligned start column      } second 16 bytes: descriptor
    float a, b;             // cos, sin of the orientation                                 }
};

// A warp owns a group of G consecutive output key points of one frame and walks three phases:
//  1. moments (IC_Angle), cooperatively per key point: lane = patch row v, the row's 31 pixels arrive as nine aligned
//     32-bit words and are reduced with dp4a against per-(alignment, |v|) weight words from shared memory (u inside the disc
//     -> m10, 1 inside the disc -> row sum -> m01 = v * row sum).  Lane k keeps the moments of key point k.
//  2. orientation, one key point per LANE: fastAtan2 and the double-precision sincos run once per 32 key points instead of
//     once per key point on all 32 lanes (a quarter of the old kernel's instructions were this redundant trigonometry).
//  3. descriptor, cooperatively per key point: the 37 x 37 blurred neighbourhood is staged in shared memory with 16-byte
//     row chunks (<= 5 coalesced loads per lane, requested one key point ahead), lane i builds byte i from 16 shared-memory
//     byte gathers.  The old kernel gathered straight from global memory: 16 warp-wide loads of 32 scattered bytes per key
//     point, ~25 L1 wavefronts each -- that, not arithmetic, was its limit.
template <int G>
__global__ void __launch_bounds__(32 * kDescWarps, 7) describe_kernel(DevPtrs d, FrameLayout fl,
                                                                  orb_keypoint_t* __restrict__ out_kps,
                                                                  uint8_t* __restrict__ out_desc,
                                                                  int32_t* __restrict__ out_counts, int cap) {
    // pattern as floats, one float4 (x0,y0,x1,y1) per test pair, laid out [pair % 8][pair / 8] so that the 32 lanes
    // (lane = descriptor byte = pair / 8) read consecutive float4s: no bank conflicts
    __shared__ float4 s_pat[256];
    __shared__ uint2 s_mw[4][16 * 9 + 1];                             // moment weights {u, 1} per (row alignment, |v| * 9 + word); the last entry is zero
    __shared__ KpWork s_kp[kDescWarps][G];
    __shared__ __align__(16) uint8_t s_patch[kDescWarps][kPatchRows * kPatchStride];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_pat[(i & 7) * 32 + (i >> 3)] = make_float4((float)c_pattern[4 * i], (float)c_pattern[4 * i + 1],
                                                     (float)c_pattern[4 * i + 2], (float)c_pattern[4 * i + 3]);
    for (int e = threadIdx.x; e < 4 * (16 * 9 + 1); e += blockDim.x) {
        const int al = e / (16 * 9 + 1), t = e % (16 * 9 + 1), av = t / 9, i = t % 9;
        uint32_t wu = 0, w1 = 0;
        for (int bb = 0; bb < 4 && t < 16 * 9; bb++) {
            const int u = 4 * i + bb - al - 15;
            if ((u < 0 ? -u : u) <= c_umax[av]) { wu |= (uint32_t)(u & 0xff) << (8 * bb); w1 |= 1u << (8 * bb); }
        }
        s_mw[al][t] = make_uint2(wu, w1);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int frame = blockIdx.y;
    // per-level slot starts / counts of this frame, one level per lane
    const int32_t* lc = d.level_count + (size_t)frame * kMaxLevels;
    const int my_cnt = lane < fl.nlevels ? lc[lane] : 0;
    const int my_slot = lane < fl.nlevels ? d.geom[lane].kp_slot : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[frame] = total;
    const int excl = incl - my_cnt;
    const int n_out = min(total, cap);
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    KpWork* const kw = s_kp[warp];
    // moment items of this lane: row offset v, byte offset of the word inside the row, byte offset of its weight pair
    int item_v[9], item_b[9], item_t[9];
#pragma unroll
    for (int it = 0; it < 9; it++) {
        const int idx = it * 32 + lane, row = idx / 9, word = idx - row * 9;
        const bool ok = row < 31;
        const int v = ok ? row - 15 : 0;
        item_v[it] = v; item_b[it] = ok ? 4 * word : 0;
        item_t[it] = 8 * (ok ? (v < 0 ? -v : v) * 9 + word : 16 * 9);
    }
    uint8_t* const patch = s_patch[warp];

    for (int first = (blockIdx.x * kDescWarps + warp) * G; first < n_out; first += gridDim.x * kDescWarps * G) {
        const int n = min(G, n_out - first);                      // key points of this group (warp-uniform)
        // ---- lane L < n: its own key point
        const int oidx = first + lane;
        int level = 0;
#pragma unroll
        for (int l = 1; l < kMaxLevels; l++) level += (l < fl.nlevels && oidx >= __shfl_sync(0xffffffffu, excl, l)) ? 1 : 0;
        const int slot = __shfl_sync(0xffffffffu, my_slot, level) + oidx - __shfl_sync(0xffffffffu, excl, level);
        uint32_t pk = 0;
        int x = 0, y = 0;
        const bool mine = lane < n;
        if (mine) {
            const LevelGeom* gp = d.geom + level;
            pk = d.sel[(size_t)frame * fl.kp_cap + slot];
            x = (int)(pk & 0xfff) + kMinBorder; y = (int)((pk >> 12) & 0xfff) + kMinBorder;
            int pitch;
            const uint8_t* roi = level_roi(d, fl, *gp, level, frame, pitch);
            const int spitch = gp->spitch;
            const int xs = (x - kPatchR) & ~15;
            KpWork w;
            w.img = roi + (size_t)y * pitch + (x - 15);       // key points sit >= 19 px inside the level, the patch radius is 15
            w.pitch = pitch;
            w.cx = (x - xs) | (spitch << 8);                  // spitch <= 4096 + 15
            w.blr = d.blur + (size_t)frame * fl.splane_bytes + gp->splane_off + (size_t)(y - kPatchR) * spitch + xs;
            w.a = 0.f; w.b = 0.f;
            kw[lane] = w;
        }
        __syncwarp();

        // ---- phase 1: IC_Angle on the un-blurred level, one key point at a time.  The 31 rows x 9 aligned words of the disc's
        // bounding box are dealt to the lanes in row-major order (item = it * 32 + lane: a warp-wide load covers 3.5 consecutive
        // rows, ~5 cache lines; with lane = row every load touched 31 lines and the kernel sat on the L1 data pipe).
        int M01 = 0, M10 = 0;
        {
            uint32_t w[9], wn[9];
            auto request = [&](int k, uint32_t (&dst)[9], int& al) {
                const int4 q = *reinterpret_cast<const int4*>(&kw[k]);            // img (2 words), pitch, cx
                const uint8_t* p = reinterpret_cast<const uint8_t*>(((uint64_t)(uint32_t)q.y << 32) | (uint32_t)q.x);
                al = (int)(reinterpret_cast<uintptr_t>(p) & 3);                   // same for every row: pitches are multiples of 4
                p -= al;
#pragma unroll
                for (int it = 0; it < 9; it++) dst[it] = __ldg(reinterpret_cast<const uint32_t*>(p + (ptrdiff_t)(item_v[it] * q.z + item_b[it])));
            };
            int al, aln = 0;
            request(0, w, al);
            for (int k = 0; k < n; k++) {
                if (k + 1 < n) request(k + 1, wn, aln);
                const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_mw[al]);
                int m01 = 0, m10 = 0;
#pragma unroll
                for (int it = 0; it < 9; it++) {
                    const uint2 t = *reinterpret_cast<const uint2*>(tab + item_t[it]);
                    m10 = dp4a_us(w[it], t.x, m10);
                    m01 += item_v[it] * dp4a_us(w[it], t.y, 0);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
                    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                }
                if (lane == k) { M01 = m01; M10 = m10; }
#pragma unroll
                for (int it = 0; it < 9; it++) w[it] = wn[it];
                al = aln;
            }
        }

        // ---- phase 2: orientation of key point L on lane L; key point record
        if (mine) {
            const float angle = fast_atan2_deg((float)M01, (float)M10);
            const float rad = __fmul_rn(angle, factorPI);
            double sd, cd;
            sincos((double)rad, &sd, &cd);
            kw[lane].a = (float)cd; kw[lane].b = (float)sd;
            const LevelGeom* gp = d.geom + level;
            orb_keypoint_t k;
            const float fx = (float)x, fy = (float)y, sc = gp->scale;
            k.x = level ? __fmul_rn(fx, sc) : fx;   // keypoint->pt *= scale  (R21 :1095-1101)
            k.y = level ? __fmul_rn(fy, sc) : fy;
            k.size = gp->patch_size;
            k.angle = angle;
            k.response = (float)(pk >> 24);
            k.octave = level;
            k.class_id = -1;
            out_kps[(size_t)frame * cap + oidx] = k;
        }
        __syncwarp();

        // ---- phase 3: steered rBRIEF on the blurred level: lane i produces descriptor byte i of one key point at a time
        {
            // staged rows: lane handles (row, chunk) = (idx >> 2, idx & 3), idx = it * 32 + lane; a chunk is needed while it
            // starts at or left of column x + 18 (every needed chunk lies inside the row: x + 18 < w <= spitch)
            uint4 st[5];
            auto request = [&](int k) {
                const int4 q0 = *reinterpret_cast<const int4*>(&kw[k]);
                const int4 q1 = *(reinterpret_cast<const int4*>(&kw[k]) + 1);
                const uint8_t* base = reinterpret_cast<const uint8_t*>(((uint64_t)(uint32_t)q1.y << 32) | (uint32_t)q1.x);
                const int cx = q0.w & 0xff, spitch = q0.w >> 8;
                const int nchunk = ((cx + kPatchR) >> 4) + 1;
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const int idx = it * 32 + lane, r = idx >> 2, c = idx & 3;
                    st[it] = make_uint4(0, 0, 0, 0);
                    if (r < kPatchRows && c < nchunk) st[it] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * spitch + c * 16));
                }
            };
            request(0);
            for (int k = 0; k < n; k++) {
                __syncwarp();                                   // the previous key point's gathers are done
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const int idx = it * 32 + lane, r = idx >> 2, c = idx & 3;
                    if (r < kPatchRows) *reinterpret_cast<uint4*>(patch + r * kPatchStride + c * 16) = st[it];
                }
                const int4 q0 = *reinterpret_cast<const int4*>(&kw[k]);
                const int4 q1 = *(reinterpret_cast<const int4*>(&kw[k]) + 1);
                __syncwarp();
                if (k + 1 < n) request(k + 1);
                const float a = __int_as_float(q1.z), b = __int_as_float(q1.w);
                const uint8_t* ctr = patch + kPatchR * kPatchStride + (q0.w & 0xff);
                int val = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 p = s_pat[j * 32 + lane];   // x0, y0, x1, y1
                    const int r0 = round_rn(__fadd_rn(__fmul_rn(p.x, b), __fmul_rn(p.y, a)));
                    const int c0 = round_rn(__fsub_rn(__fmul_rn(p.x, a), __fmul_rn(p.y, b)));
                    const int r1 = round_rn(__fadd_rn(__fmul_rn(p.z, b), __fmul_rn(p.w, a)));
                    const int c1 = round_rn(__fsub_rn(__fmul_rn(p.z, a), __fmul_rn(p.w, b)));
                    const int t0 = ctr[r0 * kPatchStride + c0];
                    const int t1 = ctr[r1 * kPatchStride + c1];
                    val |= (t0 < t1) << j;
                }
                out_desc[((size_t)frame * cap + first + k) * 32 + lane] = (uint8_t)val;
            }
        }
        __syncwarp();                                           // kw / patch are reused by the next group
    }
}

int launch_describe(const DevPtrs& d, const FrameLayout& fl, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                    int32_t* d_counts, int cap, cudaStream_t s) {
    // Batches: groups of 32 key points per warp (the per-lane trigonometry is fully used); few frames: groups of 8, four
    // times as many warps, so a single frame's ~1000 key points still spread over the GPU (latency).
    if (n_frames >= 8) {
        const int groups = (fl.kp_cap + 31) / 32;
        describe_kernel<32><<<dim3(std::max(1, (groups + kDescWarps - 1) / kDescWarps), n_frames), 32 * kDescWarps, 0, s>>>(d, fl, d_kps, d_desc, d_counts, cap);
    } else {
        const int groups = (fl.kp_cap + 7) / 8;
        describe_kernel<8><<<dim3(std::max(1, (groups + kDescWarps - 1) / kDescWarps), n_frames), 32 * kDescWarps, 0, s>>>(d, fl, d_kps, d_desc, d_counts, cap);
    }
    return 1;
}

}  // namespace orbcuda
