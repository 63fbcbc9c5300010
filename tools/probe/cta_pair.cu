// Probe for round 2: tcgen05 cta_group::2 (CTA pair) with operands written by the CTAs themselves.
//   cluster (2,1,1); both CTAs allocate TMEM with .cta_group::2; each CTA writes its A tile (128 x K32, int8) and its
//   HALF of B (N/2 = 64 rows x K32) into its own shared memory (canonical K-major no-swizzle layout); the peer signals the
//   leader through a remote mbarrier arrive; the leader issues ONE tcgen05.mma.cta_group::2 (M = 256, N = 128, K = 32) and
//   commits with multicast to both CTAs; every CTA reads its 128 x 128 accumulator back and checks it against the CPU.
// Run under `timeout`: a wrong guess about the protocol shows up as a hang or a mismatch, not as a crash.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// A: [2 CTAs][128][32] int8, B: [128][32] uint8 (rows 0-63 -> CTA 0, 64-127 -> CTA 1), D: [256][128] int32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(const int8_t* A, const uint8_t* B, int* D) {
    __shared__ __align__(1024) unsigned char s_a[128 * 32];
    __shared__ __align__(1024) unsigned char s_b[64 * 32];
    __shared__ __align__(8) unsigned long long s_full, s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cta_rank();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_full)), "r"(2));      // one arrive per CTA
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    // operands: element (r, c) at (r/8)*2048... with K = 32 only two 16-byte chunks per row: (r/8)*256 + (c/16)*128 + (r%8)*16 + c%16
    // (LBO = 128 between the K chunks, SBO = 256 between 8-row groups for this K=32-only tile)
    for (int i = tid; i < 128 * 32; i += 128) {
        const int r = i >> 5, c = i & 31;
        s_a[(r >> 3) * 256 + (c >> 4) * 128 + (r & 7) * 16 + (c & 15)] = (unsigned char)A[((size_t)rank * 128 + r) * 32 + c];
    }
    for (int i = tid; i < 64 * 32; i += 128) {
        const int r = i >> 5, c = i & 31;
        s_b[(r >> 3) * 256 + (c >> 4) * 128 + (r & 7) * 16 + (c & 15)] = B[((size_t)rank * 64 + r) * 32 + c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    cluster_sync();            // barriers initialised and TMEM allocated in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t tmem = s_tmem;
    // every CTA tells the leader (rank 0) that its operands are in place
    if (tid == 0) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&s_full)), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
    }
    if (rank == 0 && warp == 0) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&s_full)), "r"(0) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        if (leader) {
            // K-major, LBO 128, SBO 256 for this tile
            const uint64_t da = (uint64_t)((smem_u32(s_a) >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
            const uint64_t db = (uint64_t)((smem_u32(s_b) >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // s32, A signed, N=128, M=256
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u));
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         :: "r"(smem_u32(&s_done)), "h"((unsigned short)3));
        }
        __syncwarp();
    }
    // both CTAs wait for the multicast commit on their own barrier
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&s_done)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    // thread = row (lane 32*warp + lane); 128 columns in 4 loads of 32
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(trow + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[((size_t)rank * 128 + tid) * 128 + c0 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main() {
    int8_t* hA = (int8_t*)malloc(256 * 32); uint8_t* hB = (uint8_t*)malloc(128 * 32);
    srand(7);
    for (int i = 0; i < 256 * 32; i++) hA[i] = (int8_t)((rand() % 5) - 2);
    for (int i = 0; i < 128 * 32; i++) hB[i] = (uint8_t)(rand() % 4);
    int8_t* dA; uint8_t* dB; int* dD;
    cudaMalloc(&dA, 256 * 32); cudaMalloc(&dB, 128 * 32); cudaMalloc(&dD, 256 * 128 * 4);
    cudaMemcpy(dA, hA, 256 * 32, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 128 * 32, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * 128 * 4);
    pair_kernel<<<2, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    int* hD = (int*)malloc(256 * 128 * 4);
    cudaMemcpy(hD, dD, 256 * 128 * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int m = 0; m < 256; m++)
        for (int n = 0; n < 128; n++) {
            int ref = 0;
            for (int k = 0; k < 32; k++) ref += (int)hA[m * 32 + k] * (int)hB[n * 32 + k];
            if (hD[m * 128 + n] != ref) { if (bad < 6) printf("mismatch D[%d][%d] = %d, expected %d\n", m, n, hD[m * 128 + n], ref); bad++; }
        }
    printf("mismatches: %ld of %d\n", bad, 256 * 128);
    return bad != 0;
}
