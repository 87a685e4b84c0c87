// FP64 peak microbenchmarks (roofline denominators that MEASURED_PEAKS.json does not carry):
// mode 0: dependent-chain-free DFMA stream, mode 1: DMMA m8n8k4 stream.  Registers only, no memory.
#pragma once
#include "kh_common.cuh"
#ifndef KH_HOST_EMU
__global__ void __launch_bounds__(256) kh_peak_dfma(double* out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) kh_peak_dmma(double* out, int iters, double seed) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = seed + i; c1[i] = seed - i; }
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}
// mode 2: both streams interleaved in every warp (8 DMMA + 16 DFMA per iteration): tells whether DMMA and DFMA share one pipe
__global__ void __launch_bounds__(256) kh_peak_mixed(double* out, int iters, double seed) {
    double c0[8], c1[8], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = seed + i; c1[i] = seed - i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = seed + i + threadIdx.x * 1e-9;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 * (threadIdx.x + 1);
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
            asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[2 * i]) : "d"(m), "d"(c));
            asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[2 * i + 1]) : "d"(m), "d"(c));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i] + f[2 * i] + f[2 * i + 1];
    if (s == 12345.678) out[0] = s;
}
#endif
