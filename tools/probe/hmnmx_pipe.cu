// Probe: which pipe executes half2 min/max (HMNMX2 / 3-input VHMNMX) vs 16-bit-pair integer min/max (VIMNMX3.S16x2)?
// Run under: ncu --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fp16.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__global__ void k_half(unsigned* out, unsigned seed) {
    __half2 a = __halves2half2(__int2half_rn(threadIdx.x & 255), __int2half_rn((threadIdx.x * 7) & 255));
    __half2 b = __halves2half2(__int2half_rn(seed & 255), __int2half_rn((seed >> 8) & 255));
    __half2 c = __hadd2(a, b);
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) { a = __hmax2(__hmin2(a, b), c); b = __hmin2(__hmax2(b, c), a); c = __hmax2(__hmin2(c, a), b); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = *reinterpret_cast<unsigned*>(&a) ^ *reinterpret_cast<unsigned*>(&b) ^ *reinterpret_cast<unsigned*>(&c);
}
__global__ void k_int(unsigned* out, unsigned seed) {
    unsigned a = threadIdx.x * 0x00010003u, b = seed * 0x00050001u, c = a + b;
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) { a = __vimax3_s16x2(a, b, c); b = __vimin3_s16x2(b, c, a); c = __vimax3_s16x2(c, a, b) ^ 1u; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c;
}
int main() {
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(e0); k_half<<<148 * 8, 256>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("half2 min/max: %.3f ms (%d ops/thread)\n", ms, 1024 * 16 * 6);
        cudaEventRecord(e0); k_int<<<148 * 8, 256>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("s16x2 3-input min/max: %.3f ms (%d ops/thread)\n", ms, 1024 * 16 * 3);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
