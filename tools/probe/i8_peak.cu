// Probe: what the 5th-generation tensor cores of THIS B200 sustain on kind::i8 (the denominator of matching.roofline, which
// MEASURED_PEAKS.json does not hold -- it has bf16 only).  A bare issue loop: every CTA (one per SM) puts one A tile
// (128 x K32 int8) and one B tile (256 x K32) into shared memory once, then one elected thread issues
// tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32 back to back, alternating between the two 256-column halves of TMEM, with a
// commit + wait every 64 instructions so the issue queue stays bounded.  No operand traffic, no drain: nothing but the MMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/i8_peak tools/probe/i8_peak.cu && tools/probe/i8_peak
// Prints TOP/s (2 x MACs) for a burst (~1 ms) and a sustained run (~1 s).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kM = 128, kN = 256, kK = 32, kBatch = 64;

__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, unsigned long long* cycles) {
    __shared__ __align__(1024) unsigned char s_a[kM * kK];
    __shared__ __align__(1024) unsigned char s_b[kN * kK];
    __shared__ __align__(8) unsigned long long s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    for (int i = tid; i < kM * kK; i += 128) s_a[i] = (unsigned char)((i * 7 + 3) & 3);
    for (int i = tid; i < kN * kK; i += 128) s_b[i] = (unsigned char)((i * 5 + 1) & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        // K-major, no swizzle: LBO 128 between the two 16-byte K chunks, SBO 256 between 8-row groups
        const uint64_t da = (uint64_t)((smem_u32(s_a) >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
        const uint64_t db = (uint64_t)((smem_u32(s_b) >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
        const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);   // s32 accumulators, A signed
        uint32_t parity = 0;
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll 8
            for (int j = 0; j < kBatch; j++) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                             :: "r"(tmem + (uint32_t)((j & 1) * kN)), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(j > 1)), "r"(0u));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&s_done)));
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&s_done)), "r"(parity) : "memory");
            parity ^= 1;
        }
        if (cycles) cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

static double run(int sms, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    i8_peak_kernel<<<sms, 128>>>(iters, nullptr);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 2.0 * kM * kN * kK * (double)kBatch * iters * sms;
    return ops / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run(sms, 16);                                                // warm-up
    // one M128 N256 K32 MMA = 1 048 576 MACs = 128 clocks at 8192 MAC/clk/SM: 64 of them ~ 4.2 us
    const double burst = run(sms, 256);                          // ~1 ms
    const double sustained = run(sms, 256 * 1000);               // ~1 s
    printf("{\"probe\": \"tcgen05.mma kind::i8 M128 N256 K32, cta_group::1, %d CTAs\", \"burst_tops\": %.1f, \"sustained_tops\": %.1f}\n", sms, burst, sustained);
    return 0;
}
