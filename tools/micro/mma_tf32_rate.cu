// Microbenchmark: issue rate of legacy mma.sync.m16n8k8 tf32 vs FFMA2 on sm_100a (per SM, all 4 SMSPs busy).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mma_loop(float* out, int iters) {
    float c[8][4]; unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) c[j][i] = threadIdx.x * 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma2_loop(float* out, int iters) {
    unsigned long long acc[16], w = 0x3f8000003f800000ull, x = 0x3f0000003f000000ull;
    for (int j = 0; j < 16; j++) acc[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; j++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(w), "l"(x));
    }
    unsigned long long s = 0; for (int j = 0; j < 16; j++) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4 * 2);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000; float ms;
        mma_loop<<<148, warps * 32>>>(out, 10); cudaDeviceSynchronize();
        cudaEventRecord(a); mma_loop<<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        double mmas = 148.0 * warps * iters * 8; double tf = mmas * 16 * 8 * 8 * 2 / (ms * 1e-3) / 1e12;
        printf("mma.sync m16n8k8 tf32: %2d warps/SM: %.3f ms  %.1f TFLOP/s  (%.2f cycles/MMA/SMSP at 1.965 GHz)\n", warps, ms, tf, ms * 1e-3 * 1.965e9 / (iters * 8.0 * warps / 4));
        ffma2_loop<<<148, warps * 32>>>(out, 10); cudaDeviceSynchronize();
        cudaEventRecord(a); ffma2_loop<<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        double f2 = 148.0 * warps * iters * 16; printf("FFMA2              : %2d warps/SM: %.3f ms  %.1f TFLOP/s  (%.2f cycles/FFMA2/SMSP)\n", warps, ms, f2 * 32 * 2 * 2 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (iters * 16.0 * warps / 4));
    }
    return 0;
}
