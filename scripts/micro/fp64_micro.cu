// Development aid: FP64 pipe latency / single-warp issue rate on the GPU (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dep(double *out, long long *cyc, int n, int active) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-9;
    if ((int)threadIdx.x >= active) return;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
        a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k_ilp(double *out, long long *cyc, int n, int active) {
    double a[ILP];
    for (int j = 0; j < ILP; j++) a[j] = threadIdx.x * 1e-9 + j;
    const double b = 1.0000001, c = 1e-9;
    if ((int)threadIdx.x >= active) return;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) a[j] = fma(a[j], b, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; j++) s += a[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rsqrt(double *out, long long *cyc, int n) {
    double a = threadIdx.x * 1e-9 + 2.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { a = rsqrt(a) + 1.5; a = rsqrt(a) + 1.5; a = rsqrt(a) + 1.5; a = rsqrt(a) + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds(double *out, long long *cyc, int n) {
    __shared__ double s[64];
    s[threadIdx.x] = threadIdx.x; s[threadIdx.x + 32] = 1.0;
    __syncthreads();
    double a = 0.0; int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { s[idx] = a; __syncwarp(); a = s[(idx + 1) & 31] + 1.0; __syncwarp(); }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int n = 2000;
    for (int active : {1, 8, 32}) {
        k_dep<<<1, 32>>>(out, cyc, n, active); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA chain, %2d lanes: %.2f cycles/DFMA\n", active, (double)h / (8.0 * n));
        k_ilp<8><<<1, 32>>>(out, cyc, n, active); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ILP-8 DFMA, one warp, %2d lanes: %.2f cycles/DFMA\n", active, (double)h / (32.0 * n));
        k_ilp<16><<<1, 32>>>(out, cyc, n, active); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ILP-16 DFMA, one warp, %2d lanes: %.2f cycles/DFMA\n", active, (double)h / (64.0 * n));
    }
    k_ilp<8><<<1, 128>>>(out, cyc, n, 128); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ILP-8 DFMA, 4 warps (one per sub-partition): %.2f cycles/DFMA/warp\n", (double)h / (32.0 * n));
    k_rsqrt<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent rsqrt(double)+add: %.1f cycles\n", (double)h / (4.0 * n));
    k_lds<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS -> syncwarp -> LDS -> DADD -> syncwarp round trip: %.1f cycles\n", (double)h / n);
    return 0;
}
