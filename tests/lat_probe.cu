// Developer tool: dependent-chain latencies on the B200 (cycles per op, one warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, b, c);
    }
    long long t1 = clock64();
    double m = a;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) m = m * b;
    }
    long long t2 = clock64();
    double r = fabs(m) + 2.0;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r)); r = y + 2.0; }
    }
    long long t3 = clock64();
    double s = r;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
    }
    long long t4 = clock64();
    int idx = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) idx = (int)sm[idx & 31] & 31;
    }
    long long t5 = clock64();
    double q = s + 3.0;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) q = 1.0 / q + 1.5;
    }
    long long t6 = clock64();
    double w = q;
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) w = sqrt(w) + 1.5;
    }
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = (t6 - t5); cyc[6] = (t7 - t6);
        out[0] = a + m + r + s + idx + q + w;
    }
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    probe<<<1, 32>>>(out, cyc, 1.0);
    probe<<<1, 32>>>(out, cyc, 1.0);
    long long h[8];
    cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
    printf("dep DFMA %.1f cyc | dep DMUL %.1f | rsqrt.approx+DADD %.1f | 64-bit SHFL %.1f | LDS(f64)+cvt chain %.1f | 1/x + add %.1f | sqrt + add %.1f\n",
           h[0] / 16000.0, h[1] / 16000.0, h[2] / 4000.0, h[3] / 8000.0, h[4] / 8000.0, h[5] / 4000.0, h[6] / 4000.0);
    return 0;
}
