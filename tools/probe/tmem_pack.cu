// Probe: what does tcgen05.ld ... .pack::16b return?  (writes known int32 values with tcgen05.st, reads them back packed)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t t = s_tmem + ((threadIdx.x & 96u) << 16);   // warp w -> lanes 32w
    uint32_t v[16];
    for (int c = 0; c < 16; c++) v[c] = (uint32_t)((c & 1) ? -(int)(threadIdx.x * 16 + c) : (int)(threadIdx.x * 16 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(t), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(t));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) out[threadIdx.x * 8 + i] = r[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "r"(64));
}
int main() {
    uint32_t* d; cudaMalloc(&d, 128 * 8 * 4);
    probe<<<1, 128>>>(d);
    uint32_t h[128 * 8];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    for (int th : {0, 1, 37}) {
        printf("thread %d:", th);
        for (int i = 0; i < 8; i++) printf(" [lo %d hi %d]", (int)(int16_t)(h[th * 8 + i] & 0xffff), (int)(int16_t)(h[th * 8 + i] >> 16));
        printf("\n");
    }
    return 0;
}
