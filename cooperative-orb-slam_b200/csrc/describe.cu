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

__global__ void __launch_bounds__(256) describe_kernel(DevPtrs d, FrameLayout fl, orb_keypoint_t* __restrict__ out_kps,
                                                       uint8_t* __restrict__ out_desc, int32_t* __restrict__ out_counts,
                                                       int cap) {
    __shared__ int8_t s_pat[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_pat[i] = c_pattern[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // key point slot inside the frame
    const int frame = blockIdx.y;
    if (slot >= fl.kp_cap) return;
    // which level / rank does this slot belong to, and where does it land in the level-major output
    const int32_t* lc = d.level_count + (size_t)frame * kMaxLevels;
    int level = 0, prefix = 0, total = 0;
    for (int l = 0; l < fl.nlevels; l++) {
        const int c = lc[l];
        if (slot >= d.geom[l].kp_slot) { level = l; prefix = total; }
        total += c;
    }
    if (slot == 0 && lane == 0) out_counts[frame] = total;
    const LevelGeom g = d.geom[level];
    const int rank = slot - g.kp_slot;
    if (rank >= lc[level]) return;
    const int oidx = prefix + rank;
    if (oidx >= cap) return;
    const uint32_t pk = d.sel[(size_t)frame * fl.kp_cap + slot];
    const int x = (int)(pk & 0xfff) + kMinBorder, y = (int)((pk >> 12) & 0xfff) + kMinBorder;
    const int score = (int)(pk >> 24);

    // ---- IC_Angle on the un-blurred padded level: lanes sweep u = -15..15, rows v = -15..15
    const uint8_t* img = d.pyr + (size_t)frame * fl.pyr_bytes + g.plane_off + (size_t)(kEdge + y) * g.pitch + kXPad + x;
    int m01 = 0, m10 = 0;
    const int u = lane - 15;
    if (lane < 31) {
#pragma unroll 1
        for (int v = -15; v <= 15; v++) {
            const int lim = c_umax[v < 0 ? -v : v];
            if (u >= -lim && u <= lim) {
                const int p = img[(ptrdiff_t)v * g.pitch + u];
                m10 += u * p;
                m01 += v * p;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered rBRIEF on the blurred level: lane i produces descriptor byte i
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const uint8_t* blr = d.blur + (size_t)frame * fl.splane_bytes + g.splane_off + (size_t)y * g.spitch + x;
    const int8_t* pat = s_pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1];
        const float x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = blr[(ptrdiff_t)r0 * g.spitch + c0];
        const int t1 = blr[(ptrdiff_t)r1 * g.spitch + c1];
        val |= (t0 < t1) << j;
    }
    out_desc[((size_t)frame * cap + oidx) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orb_keypoint_t k;
        const float fx = (float)x, fy = (float)y;
        k.x = level ? __fmul_rn(fx, g.scale) : fx;   // keypoint->pt *= scale  (R21 :1095-1101)
        k.y = level ? __fmul_rn(fy, g.scale) : fy;
        k.size = g.patch_size;
        k.angle = angle;
        k.response = (float)score;
        k.octave = level;
        k.class_id = -1;
        out_kps[(size_t)frame * cap + oidx] = k;
    }
}

int launch_describe(const DevPtrs& d, const FrameLayout& fl, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                    int32_t* d_counts, int cap, cudaStream_t s) {
    const int threads = 256;
    const int blocks = (fl.kp_cap * 32 + threads - 1) / threads;
    describe_kernel<<<dim3(blocks, n_frames), threads, 0, s>>>(d, fl, d_kps, d_desc, d_counts, cap);
    return 1;
}

}  // namespace orbcuda
