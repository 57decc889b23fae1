#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void k_dep(double *out, long long *cyc, int n) {
    double a = threadIdx.x * 1e-3 + 0.5, b = 0.25, c0 = 0, c1 = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        dmma(c0, c1, a, b, c0, c1); dmma(c0, c1, a, b, c0, c1); dmma(c0, c1, a, b, c0, c1); dmma(c0, c1, a, b, c0, c1);
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// chain through the A operand (result feeds next A): the Riccati pattern
__global__ void k_depA(double *out, long long *cyc, int n) {
    double a = threadIdx.x * 1e-3 + 0.5, b = 0.25, c0 = 0, c1 = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        dmma(c0, c1, a, b, 0.0, 0.0); a = c0; dmma(c0, c1, a, b, 0.0, 0.0); a = c1;
        dmma(c0, c1, a, b, 0.0, 0.0); a = c0; dmma(c0, c1, a, b, 0.0, 0.0); a = c1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1 + a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ilp(double *out, long long *cyc, int n) {
    double a = threadIdx.x * 1e-3 + 0.5, b = 0.25, c[16];
    for (int j = 0; j < 16; j++) c[j] = j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(c[2*j], c[2*j+1], a, b, c[2*j], c[2*j+1]);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 16; j++) s += c[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc, int n) {
    double a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31) + 1.0; a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 3) & 31) + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int n = 2000;
    k_dep<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DMMA m8n8k4 chain (accumulator): %.2f cycles/DMMA\n", (double)h / (4.0 * n));
    k_depA<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DMMA chain (through A operand): %.2f cycles/DMMA\n", (double)h / (4.0 * n));
    k_ilp<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("independent DMMA x8, one warp: %.2f cycles/DMMA\n", (double)h / (8.0 * n));
    k_ilp<<<1, 128>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("independent DMMA x8, 4 warps: %.2f cycles/DMMA/warp\n", (double)h / (8.0 * n));
    k_shfl<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("double shfl + DADD dependent: %.2f cycles\n", (double)h / (2.0 * n));
    return 0;
}
