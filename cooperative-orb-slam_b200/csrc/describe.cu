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

constexpr int kDescWarps = 8;

__global__ void __launch_bounds__(32 * kDescWarps) describe_kernel(DevPtrs d, FrameLayout fl,
                                                                  orb_keypoint_t* __restrict__ out_kps,
                                                                  uint8_t* __restrict__ out_desc,
                                                                  int32_t* __restrict__ out_counts, int cap) {
    // pattern as floats, one float4 (x0,y0,x1,y1) per test pair, laid out [pair % 8][pair / 8] so that the 32 lanes
    // (lane = descriptor byte = pair / 8) read consecutive float4s: no bank conflicts
    __shared__ float4 s_pat[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_pat[(i & 7) * 32 + (i >> 3)] = make_float4((float)c_pattern[4 * i], (float)c_pattern[4 * i + 1],
                                                     (float)c_pattern[4 * i + 2], (float)c_pattern[4 * i + 3]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int frame = blockIdx.y;
    // per-level slot starts / counts of this frame, one level per lane
    const int32_t* lc = d.level_count + (size_t)frame * kMaxLevels;
    const int my_cnt = lane < fl.nlevels ? lc[lane] : 0;
    const int my_slot = lane < fl.nlevels ? d.geom[lane].kp_slot : 0x7fffffff;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[frame] = total;
    const float factorPI = (float)(3.14159265358979323846 / 180.f);

    for (int slot = blockIdx.x * kDescWarps + warp; slot < fl.kp_cap; slot += gridDim.x * kDescWarps) {
        const int level = __popc(__ballot_sync(0xffffffffu, slot >= my_slot)) - 1;
        const int rank = slot - __shfl_sync(0xffffffffu, my_slot, level);
        const int cnt = __shfl_sync(0xffffffffu, my_cnt, level);
        const int oidx = __shfl_sync(0xffffffffu, incl - my_cnt, level) + rank;
        if (rank >= cnt || oidx >= cap) continue;
        const LevelGeom* gp = d.geom + level;
        const int spitch = gp->spitch;
        int pitch;
        const uint8_t* roi = level_roi(d, fl, *gp, level, frame, pitch);
        const uint32_t pk = d.sel[(size_t)frame * fl.kp_cap + slot];
        const int x = (int)(pk & 0xfff) + kMinBorder, y = (int)((pk >> 12) & 0xfff) + kMinBorder;

        // ---- IC_Angle on the un-blurred padded level.  Lane u sweeps its column of the disc:
        // |v| <= umax[|u|] (the patch is symmetric by construction, :462-469).  All 31 predicated loads are
        // issued back to back (fully unrolled) so their latencies overlap.
        const uint8_t* img = roi + (size_t)y * pitch + x;   // key points sit >= 19 px inside the level, the patch radius is 15
        int m01 = 0, m10 = 0;
        {
            const int u = lane - 15;
            const int au = u < 0 ? -u : u;
            const int R = lane < 31 ? c_umax[au] : -1;
            int pv[31];
#pragma unroll
            for (int i = 0; i < 31; i++) {
                const int v = i - 15;
                const int av = v < 0 ? -v : v;
                pv[i] = av <= R ? (int)img[(ptrdiff_t)v * pitch + u] : 0;
            }
            int colsum = 0;
#pragma unroll
            for (int i = 0; i < 31; i++) {
                colsum += pv[i];
                m01 += (i - 15) * pv[i];
            }
            m10 = u * colsum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- steered rBRIEF on the blurred level: lane i produces descriptor byte i
        const float rad = __fmul_rn(angle, factorPI);
        double sd, cd;
        sincos((double)rad, &sd, &cd);
        const float a = (float)cd, b = (float)sd;
        const uint8_t* blr = d.blur + (size_t)frame * fl.splane_bytes + gp->splane_off + (size_t)y * spitch + x;
        int val = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 p = s_pat[j * 32 + lane];   // x0, y0, x1, y1
            const int r0 = round_rn(__fadd_rn(__fmul_rn(p.x, b), __fmul_rn(p.y, a)));
            const int c0 = round_rn(__fsub_rn(__fmul_rn(p.x, a), __fmul_rn(p.y, b)));
            const int r1 = round_rn(__fadd_rn(__fmul_rn(p.z, b), __fmul_rn(p.w, a)));
            const int c1 = round_rn(__fsub_rn(__fmul_rn(p.z, a), __fmul_rn(p.w, b)));
            const int t0 = blr[r0 * spitch + c0];
            const int t1 = blr[r1 * spitch + c1];
            val |= (t0 < t1) << j;
        }
        out_desc[((size_t)frame * cap + oidx) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
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
    }
}

int launch_describe(const DevPtrs& d, const FrameLayout& fl, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                    int32_t* d_counts, int cap, cudaStream_t s) {
    // each warp walks ~4 key point slots, amortising the pattern staging
    const int blocks = std::max(1, (fl.kp_cap + kDescWarps * 4 - 1) / (kDescWarps * 4));
    describe_kernel<<<dim3(blocks, n_frames), 32 * kDescWarps, 0, s>>>(d, fl, d_kps, d_desc, d_counts, cap);
    return 1;
}

}  // namespace orbcuda
