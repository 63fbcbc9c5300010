// describe.cu -- K5: IC_Angle orientation + steered rBRIEF descriptor, fused; persistent warps, a warp owns a group of key
// points, both pixel neighbourhoods of a key point arrive in shared memory by TMA bulk tensor copies
// (replaces computeOrientation/IC_Angle R21/src/ORBextractor.cc:472-479,:77-104 and
// computeDescriptors/computeOrbDescriptor :1034-1041,:108-147, plus the level -> image coordinate
// scaling and level-major concatenation of operator() :1072-1104).
//
// Float parity: fastAtan2 is OpenCV's degree-7 polynomial evaluated in strict fp32 with separate
// multiplies and adds (__fmul_rn/__fadd_rn: no FMA contraction); the pattern rotation is
// x*b + y*a / x*a - y*b again without FMA and rounded with cvRound == round-half-even
// (magic-number add).  cos/sin are evaluated in double and rounded once to float (the reference calls
// libm cosf/sinf, which is not correctly rounded and differs between glibc builds).
#include "internal.h"
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

namespace orbcuda {

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

constexpr int kPatchR = 18;                     // the pattern's largest radius is 18.38 (point (-13,-13)): |cvRound(rotated coordinate)| <= 18
constexpr int kBlurBoxW = 80, kBlurBoxH = 2 * kPatchR + 1;    // TMA box of the blurred level: 37 rows x 80 bytes from the 16-byte aligned column at or left of x - 18
constexpr int kImgBoxW = 48, kImgBoxH = 31;                   // TMA box of the un-blurred level: 31 rows x 48 bytes from the aligned column at or left of x - 15
constexpr int kStageBytes = 3072;               // one staging buffer (>= 80 * 37 = 2960, a multiple of 128)

__device__ __forceinline__ int dp4a_us(uint32_t px, uint32_t w, int acc) {   // 4 unsigned pixels x 4 signed weights
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(px), "r"(w), "r"(acc));
    return r;
}
__device__ __forceinline__ uint32_t dsc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dsc_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// one elected lane: arm the barrier with the box size and start the bulk tensor copy (TMA) of one box into shared memory;
// columns / rows of the box outside the tensor are zero-filled by the copy engine (they are never used)
__device__ __forceinline__ void dsc_tma_box(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int frame, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(frame) : "memory");
}

// ---- compile-time tables (copied to shared memory by every CTA with coalesced loads) ----
constexpr int kUmaxHost[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
constexpr int8_t kPatternHost[1024] = {
#include "orb_pattern.inc"
};
constexpr uint32_t half_bits_of_small_int(int v) {     // fp16 bit pattern of an integer |v| < 2048
    if (v == 0) return 0;
    const uint32_t sign = v < 0 ? 0x8000u : 0u;
    uint32_t a = (uint32_t)(v < 0 ? -v : v);
    int e = 0;
    while ((a >> (e + 1)) != 0) e++;
    return sign | ((uint32_t)(e + 15) << 10) | ((a << (10 - e)) & 0x3ffu);
}
// test pattern as fp16 pairs {(x0, y0), (x1, y1)} per pair, in pair order (the coordinates are small integers: exact)
struct PatTable { uint2 v[256]; };
constexpr PatTable make_pat_table() {
    PatTable t{};
    for (int i = 0; i < 256; i++) {
        t.v[i].x = half_bits_of_small_int(kPatternHost[4 * i]) | (half_bits_of_small_int(kPatternHost[4 * i + 1]) << 16);
        t.v[i].y = half_bits_of_small_int(kPatternHost[4 * i + 2]) | (half_bits_of_small_int(kPatternHost[4 * i + 3]) << 16);
    }
    return t;
}
// moment weights: entry [al][v + 15][word] = {u, v} for the four pixels of an aligned word of patch row v when the row's first
// pixel (u = -15) sits at byte `al` of word 0; zero outside the disc |u| <= umax[|v|] (R21 :462-469, :77-104)
constexpr int kMwEntries = 4 * 31 * 9;
struct MwTable { uint2 v[kMwEntries]; };
constexpr MwTable make_mw_table() {
    MwTable t{};
    for (int al = 0; al < 4; al++)
        for (int row = 0; row < 31; row++)
            for (int w = 0; w < 9; w++) {
                const int v = row - 15, av = v < 0 ? -v : v;
                uint32_t wu = 0, wv = 0;
                for (int bb = 0; bb < 4; bb++) {
                    const int u = 4 * w + bb - al - 15;
                    if ((u < 0 ? -u : u) <= kUmaxHost[av]) { wu |= (uint32_t)(u & 0xff) << (8 * bb); wv |= (uint32_t)(v & 0xff) << (8 * bb); }
                }
                t.v[(al * 31 + row) * 9 + w].x = wu;
                t.v[(al * 31 + row) * 9 + w].y = wv;
            }
    return t;
}
__device__ const PatTable g_pat = make_pat_table();
__device__ const MwTable g_mw = make_mw_table();

struct TmapSet { CUtensorMap img[kMaxLevels]; CUtensorMap blr[kMaxLevels]; };   // [B][rows][pitch] uint8 tensors of every level

// A warp owns a group of G consecutive output key points of one frame and walks three phases:
//  1. moments (IC_Angle), cooperatively per key point: the 31 x 48-byte neighbourhood of the un-blurred level arrives in
//     shared memory as ONE bulk tensor copy (TMA, requested one key point ahead, completion on an mbarrier); lane = (row mod 3,
//     word) reads one aligned word of 11 rows and reduces it with two dp4a against per-(alignment, row, word) weight words
//     (u inside the disc -> m10, v inside the disc -> m01).  Lane k keeps the moments of key point k.
//  2. orientation, one key point per LANE: fastAtan2 and the double-precision sincos run once per 32 key points instead of
//     once per key point on all 32 lanes (a quarter of the first kernel's instructions were this redundant trigonometry).
//  3. descriptor, cooperatively per key point: the 37 x 80-byte blurred neighbourhood arrives by TMA the same way; 8 rounds of
//     32 test pairs, lane L evaluates pair 32 j + L from two shared-memory byte gathers, the warp ballot of round j is
//     descriptor word j.  The first kernel gathered straight from global memory: 16 warp-wide loads of 32 scattered bytes
//     per key point, ~25 L1 wavefronts each -- that, not arithmetic, was its limit.
template <int G, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) describe_kernel(const __grid_constant__ TmapSet tm, DevPtrs d, FrameLayout fl,
                                                                    orb_keypoint_t* __restrict__ out_kps,
                                                                    uint8_t* __restrict__ out_desc,
                                                                    int32_t* __restrict__ out_counts, int cap, int n_frames, int groups_per_frame) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    // layout: [WARPS][2][kStageBytes] staging | s_mw | s_pat | [WARPS][G] key point records | [WARPS][2] barriers
    uint8_t* const s_stage = s_dyn;
    uint2* const s_mw = reinterpret_cast<uint2*>(s_dyn + WARPS * 2 * kStageBytes);
    uint2* const s_pat = s_mw + kMwEntries;
    int4* const s_kp = reinterpret_cast<int4*>(s_pat + 256);                 // x | y << 16, level, cos, sin
    uint64_t* const s_bar = reinterpret_cast<uint64_t*>(s_kp + WARPS * G);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_pat[i] = g_pat.v[i];
    for (int i = threadIdx.x; i < kMwEntries; i += blockDim.x) s_mw[i] = g_mw.v[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(dsc_smem_u32(&s_bar[2 * warp])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(dsc_smem_u32(&s_bar[2 * warp + 1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    __syncthreads();
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    int4* const kw = s_kp + warp * G;
    const uint32_t stage0 = dsc_smem_u32(s_stage + (size_t)warp * 2 * kStageBytes);
    const uint32_t bar0 = dsc_smem_u32(&s_bar[2 * warp]);
    uint32_t par = 0;                                            // phase parity of the two barriers (bit b = barrier b)

    // Persistent warps: work item = (frame, group of G output key points); every warp of the grid strides over the items, so the
    // load is even whatever the batch size (one launch wave; with one CTA per (frame, 4 groups) the last wave ran a tenth full).
    const int n_items = n_frames * groups_per_frame;
    for (int item = blockIdx.x * WARPS + warp; item < n_items; item += gridDim.x * WARPS) {
    const int frame = item / groups_per_frame;
    const int first = (item - frame * groups_per_frame) * G;
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
    if (first == 0 && lane == 0) out_counts[frame] = total;
    const int excl = incl - my_cnt;
    const int n_out = min(total, cap);
    if (first >= n_out) continue;
    {
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
            pk = d.sel[(size_t)frame * fl.kp_cap + slot];
            x = (int)(pk & 0xfff) + kMinBorder; y = (int)((pk >> 12) & 0xfff) + kMinBorder;
            kw[lane] = make_int4(x | (y << 16), level, 0, 0);
        }
        __syncwarp();

        // ---- phase 1: IC_Angle on the un-blurred level, one key point at a time (key points sit >= 19 px inside the level, the
        // patch radius is 15).  Levels >= 1 live in padded planes: pixel (x, y) is element (kXPad + x, kEdge + y) of the tensor.
        int M01 = 0, M10 = 0;
        {
            const int lr = lane / 9, lw = lane - 9 * lr;          // lanes 27..31 repeat rows of other lanes: zeroed below
            const uint32_t lane_px = stage0 + lr * kImgBoxW + 4 * lw;
            const unsigned char* const lane_tab = reinterpret_cast<const unsigned char*>(s_mw) + (lr * 9 + lw) * 8;
            auto request = [&](int k) {
                const int4 q = kw[k];
                const int kx = q.x & 0xffff, ky = q.x >> 16, lv = q.y;
                dsc_tma_box(stage0 + (k & 1) * kStageBytes, &tm.img[lv], bar0 + 8 * (k & 1), ((kx - 15) & ~15) + (lv ? kXPad : 0),
                            ky - 15 + (lv ? kEdge : 0), frame, kImgBoxW * kImgBoxH);
            };
            if (lane == 0) request(0);
            for (int k = 0; k < n; k++) {
                const int cur = k & 1;
                if (lane == 0 && k + 1 < n) request(k + 1);       // its buffer was last read in iteration k - 1 (warp barrier below)
                const int al16 = ((kw[k].x & 0xffff) - 15) & 15;
                const uint32_t px = lane_px + cur * kStageBytes + (al16 & ~3);
                const unsigned char* tab = lane_tab + (al16 & 3) * (31 * 9 * 8);
                dsc_mbar_wait(bar0 + 8 * cur, (par >> cur) & 1);
                par ^= 1u << cur;
                int m01 = 0, m10 = 0;
#pragma unroll
                for (int it = 0; it < 11; it++) {
                    if (it < 10 || lr == 0) {                     // row 3 it + lr < 31
                        uint32_t w;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(px + it * 3 * kImgBoxW));
                        const uint2 t = *reinterpret_cast<const uint2*>(tab + it * 3 * 9 * 8);
                        m10 = dp4a_us(w, t.x, m10);
                        m01 = dp4a_us(w, t.y, m01);
                    }
                }
                if (lane >= 27) { m01 = 0; m10 = 0; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
                    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                }
                if (lane == k) { M01 = m01; M10 = m10; }
                __syncwarp();
            }
        }

        // ---- phase 2: orientation of key point L on lane L; key point record
        if (mine) {
            const float angle = fast_atan2_deg((float)M01, (float)M10);
            const float rad = __fmul_rn(angle, factorPI);
            double sd, cd;
            sincos((double)rad, &sd, &cd);
            kw[lane].z = __float_as_int((float)cd); kw[lane].w = __float_as_int((float)sd);
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

        // ---- phase 3: steered rBRIEF on the blurred level, one key point at a time: 8 rounds of 32 test pairs
        {
            // this lane's test pairs in registers for the 32 key points of the group: round j, lane L evaluates pair 32 j + L
            uint2 pat[8];
#pragma unroll
            for (int j = 0; j < 8; j++) pat[j] = s_pat[32 * j + lane];
            auto request = [&](int k) {
                const int4 q = kw[k];
                const int kx = q.x & 0xffff, ky = q.x >> 16;
                dsc_tma_box(stage0 + (k & 1) * kStageBytes, &tm.blr[q.y], bar0 + 8 * (k & 1), (kx - kPatchR) & ~15, ky - kPatchR, frame,
                            kBlurBoxW * kBlurBoxH);
            };
            if (lane == 0) request(0);
            for (int k = 0; k < n; k++) {
                const int cur = k & 1;
                if (lane == 0 && k + 1 < n) request(k + 1);
                const int4 q = kw[k];
                const float a = __int_as_float(q.z), b = __int_as_float(q.w);
                // byte (r, c) of the patch is at stage + (18 + r) * 80 + cx + c; the rounding constant 0x4b400000 that round_rn
                // would subtract from each coordinate is folded into the base (32-bit shared addresses wrap)
                const uint32_t ctr = stage0 + cur * kStageBytes + kPatchR * kBlurBoxW + (uint32_t)(((q.x & 0xffff) - kPatchR) & 15) + kPatchR
                                     - 0x4b400000u * (uint32_t)(kBlurBoxW + 1);
                dsc_mbar_wait(bar0 + 8 * cur, (par >> cur) & 1);
                par ^= 1u << cur;
                uint32_t word = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&pat[j].x));   // (x0, y0)
                    const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&pat[j].y));   // (x1, y1)
                    // cvRound(x*b + y*a), cvRound(x*a - y*b): adding 1.5 * 2^23 leaves rint() in the low mantissa bits
                    const uint32_t r0 = (uint32_t)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)), 12582912.f));
                    const uint32_t c0 = (uint32_t)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)), 12582912.f));
                    const uint32_t r1 = (uint32_t)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a)), 12582912.f));
                    const uint32_t c1 = (uint32_t)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b)), 12582912.f));
                    uint32_t t0, t1;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(ctr + r0 * kBlurBoxW + c0));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(ctr + r1 * kBlurBoxW + c1));
                    const uint32_t bits = __ballot_sync(0xffffffffu, t0 < t1);
                    if (lane == j) word = bits;
                }
                if (lane < 8) reinterpret_cast<uint32_t*>(out_desc + ((size_t)frame * cap + first + k) * 32)[lane] = word;
                __syncwarp();                                     // every lane is done with this buffer before it is refilled
            }
        }
    }
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct DescribeMaps {
    TmapSet set;
    // what the maps were built for
    const void* in = nullptr; const void* pyr = nullptr; const void* blur = nullptr;
    size_t in_stride = 0; int in_pitch = 0, width = 0, height = 0, nlevels = 0, frames = 0;
    int64_t pyr_bytes = 0, splane_bytes = 0;
    bool valid = false;
};
DescribeMaps* describe_maps_create() { return new DescribeMaps(); }
void describe_maps_destroy(DescribeMaps* m) { delete m; }

static bool encode_u8_3d(CUtensorMap* map, const void* base, uint64_t pitch, uint64_t rows, uint64_t frames, uint64_t frame_stride, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from this driver"); return false; }
    const cuuint64_t dims[3] = {pitch, rows, frames};
    const cuuint64_t strides[2] = {pitch, frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): base %p pitch %llu rows %llu frames %llu stride %llu", (int)r, base,
                                       (unsigned long long)pitch, (unsigned long long)rows, (unsigned long long)frames, (unsigned long long)frame_stride); return false; }
    return true;
}

constexpr int kDescWarps = 8;
template <int G, int WARPS> constexpr size_t describe_smem_bytes() {
    return (size_t)WARPS * 2 * kStageBytes + (size_t)(kMwEntries + 256) * sizeof(uint2) + (size_t)WARPS * G * sizeof(int4) + (size_t)WARPS * 2 * sizeof(uint64_t);
}
struct DescribeGrid { int blocks_per_sm = 0, sms = 0; };
template <int G, int WARPS, int MINB>
static int launch_describe_g(const TmapSet& set, const DevPtrs& d, const FrameLayout& fl, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                             int32_t* d_counts, int cap, cudaStream_t s) {
    constexpr size_t smem = describe_smem_bytes<G, WARPS>();
    static DeviceOnce once;
    static DescribeGrid grid[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!once.run([&] {
            if (cudaFuncSetAttribute(describe_kernel<G, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
            int bps = 0, sms = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, describe_kernel<G, WARPS, MINB>, 32 * WARPS, smem) != cudaSuccess || bps < 1) return false;
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
            grid[dev].blocks_per_sm = bps; grid[dev].sms = sms;
            return true;
        })) { set_error("describe kernel configuration failed"); return -1; }
    const int gpf = (fl.kp_cap + G - 1) / G;
    const int items = n_frames * gpf;
    const int blocks = std::max(1, std::min((items + WARPS - 1) / WARPS, grid[dev].blocks_per_sm * grid[dev].sms));
    describe_kernel<G, WARPS, MINB><<<blocks, 32 * WARPS, smem, s>>>(set, d, fl, d_kps, d_desc, d_counts, cap, n_frames, gpf);
    return 1;
}
static int launch_describe_variant(const TmapSet& set, const DevPtrs& d, const FrameLayout& fl, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                                   int32_t* d_counts, int cap, cudaStream_t s) {
    // Groups of G key points per work item: the per-lane trigonometry of a group costs the same for any G, small groups
    // balance better and spread a single frame over the whole GPU (measured on 128-frame batches: 0.099 ms with G = 8, 0.106
    // with 16, 0.132 with 32; single frame: 17 us with G = 8, 23 with 16, 38 with 32).  ORBCUDA_DESC_G = 4 | 8 | 16 | 32 forces one.
    static const int forced = [] { const char* e = getenv("ORBCUDA_DESC_G"); return e ? atoi(e) : 0; }();
    const int g = forced ? forced : (n_frames < 4 ? 4 : 8);
    if (g >= 32) return launch_describe_g<32, kDescWarps, 3>(set, d, fl, n_frames, d_kps, d_desc, d_counts, cap, s);
    if (g >= 16) return launch_describe_g<16, kDescWarps, 3>(set, d, fl, n_frames, d_kps, d_desc, d_counts, cap, s);
    if (g >= 8) return launch_describe_g<8, kDescWarps, 3>(set, d, fl, n_frames, d_kps, d_desc, d_counts, cap, s);
    return launch_describe_g<4, kDescWarps, 3>(set, d, fl, n_frames, d_kps, d_desc, d_counts, cap, s);
}

int launch_describe(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hg, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                    int32_t* d_counts, int cap, DescribeMaps* m, cudaStream_t s) {
    // the tensor maps only change with the buffers / the input pointer: rebuilt on demand (16 host-side encodes)
    if (!m->valid || m->in != d.in || m->pyr != d.pyr || m->blur != d.blur || m->in_stride != d.in_frame_stride || m->in_pitch != fl.in_pitch ||
        m->width != fl.width || m->height != fl.height || m->nlevels != fl.nlevels || m->frames < n_frames || m->pyr_bytes != fl.pyr_bytes ||
        m->splane_bytes != fl.splane_bytes) {
        m->valid = false;
        for (int l = 0; l < fl.nlevels; l++) {
            const bool ok_img = l == 0 ? encode_u8_3d(&m->set.img[0], d.in, (uint64_t)fl.in_pitch, (uint64_t)fl.height, (uint64_t)n_frames, d.in_frame_stride, kImgBoxW, kImgBoxH)
                                       : encode_u8_3d(&m->set.img[l], d.pyr + hg[l].plane_off, (uint64_t)hg[l].pitch, (uint64_t)hg[l].plane_rows, (uint64_t)n_frames,
                                                      (uint64_t)fl.pyr_bytes, kImgBoxW, kImgBoxH);
            if (!ok_img) return -1;
            if (!encode_u8_3d(&m->set.blr[l], d.blur + hg[l].splane_off, (uint64_t)hg[l].spitch, (uint64_t)hg[l].h, (uint64_t)n_frames, (uint64_t)fl.splane_bytes,
                              kBlurBoxW, kBlurBoxH)) return -1;
        }
        for (int l = fl.nlevels; l < kMaxLevels; l++) { m->set.img[l] = m->set.img[0]; m->set.blr[l] = m->set.blr[0]; }
        m->in = d.in; m->pyr = d.pyr; m->blur = d.blur; m->in_stride = d.in_frame_stride; m->in_pitch = fl.in_pitch; m->width = fl.width;
        m->height = fl.height; m->nlevels = fl.nlevels; m->frames = n_frames; m->pyr_bytes = fl.pyr_bytes; m->splane_bytes = fl.splane_bytes;
        m->valid = true;
    }
    return launch_describe_variant(m->set, d, fl, n_frames, d_kps, d_desc, d_counts, cap, s);
}

}  // namespace orbcuda
