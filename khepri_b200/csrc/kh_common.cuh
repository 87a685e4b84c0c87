// khepri_b200 -- common definitions for the sm_100a RCWA kernels.
//
// All kernels are written against a tiny "CTA context" (struct Cta) instead of raw threadIdx /
// __syncthreads so that the very same source also compiles as plain host C++ with
// -DKH_HOST_EMU (one virtual thread per CTA, barriers are no-ops).  The host-emulation build is
// TEST INFRASTRUCTURE (tests/hostemu): it lets the algorithm logic (eigensolver, Gauss-Jordan,
// star-product orchestration) be checked on a machine without a GPU.  The product library is
// only ever built by nvcc for sm_100a and has no CPU fallback.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef KH_HOST_EMU
#define KH_HD inline
#define KH_DEV inline
#include <algorithm>
typedef void* kh_stream_t;
#else
#include <cuda_runtime.h>
#define KH_HD __host__ __device__ __forceinline__
#define KH_DEV __device__ __forceinline__
typedef cudaStream_t kh_stream_t;
#endif

// ------------------------------------------------------------------ complex128
struct alignas(16) cd {
    double x, y;
};
KH_HD cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
KH_HD cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
KH_HD cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
KH_HD cd operator-(cd a) { return mk(-a.x, -a.y); }
KH_HD cd operator*(cd a, cd b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
KH_HD cd operator*(double s, cd a) { return mk(s * a.x, s * a.y); }
KH_HD cd operator*(cd a, double s) { return mk(s * a.x, s * a.y); }
KH_HD cd cconj(cd a) { return mk(a.x, -a.y); }
KH_HD double cabs1(cd a) { return fabs(a.x) + fabs(a.y); }
KH_HD double cabs2(cd a) { return a.x * a.x + a.y * a.y; }
KH_HD double cabsd(cd a) { return hypot(a.x, a.y); }
// c += a*b
KH_HD void cfma(cd& c, cd a, cd b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
// write-once output streams (field maps): evict-first store, the data is not read again by the kernel
KH_HD void kh_store_stream(cd* p, cd v) {
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
#else
    *p = v;
#endif
}
// c -= a*b
KH_HD void cfms(cd& c, cd a, cd b) {
    c.x = fma(-a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
}
// Smith's algorithm (what numpy uses for complex division)
KH_HD cd operator/(cd a, cd b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
KH_HD cd crecip(cd b) { return mk(1.0, 0.0) / b; }
// principal square root with C99 signed-zero semantics (matches numpy.sqrt on complex input)
KH_HD cd csqrt_(cd z) {
    if (z.x == 0.0 && z.y == 0.0) return mk(0.0, z.y);
    double t = sqrt(0.5 * (fabs(z.x) + hypot(z.x, z.y)));
    if (z.x >= 0.0) return mk(t, z.y / (2.0 * t));
    return mk(fabs(z.y) / (2.0 * t), copysign(t, z.y));
}
KH_HD cd cexp_(cd z) {
    double e = exp(z.x), s, c;
#ifdef __CUDA_ARCH__
    sincos(z.y, &s, &c);
#else
    s = sin(z.y); c = cos(z.y);
#endif
    return mk(e * c, e * s);
}

// 1/sqrt(x) for normal positive x: hardware seed (MUFU.RSQ64H) + two Newton steps, branch free.
// Used on the critical path of the Givens rotations (library rsqrt()/sqrt()/division carry slow-path branches).
KH_HD double kh_rsqrt(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double hx = 0.5 * x;                      // seed ~ 2^-20 relative: two Newton steps reach double precision
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}

#ifdef __CUDACC__
// Division-free scalar helpers for latency-critical, redundantly computed scalars (Wilkinson shifts, Householder scalars):
// on B200 a dependent FP64 op costs ~20 cycles and the library sqrt / division / hypot are 25-40 of them back to back.
// 1/x for normal x: hardware seed (20 bits) + two Newton steps.
__device__ __forceinline__ double kh_rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
// principal complex square root, ~2 ulp: exact power-of-four scaling, two reciprocal square roots, no division / hypot.
__device__ __forceinline__ cd kh_csqrt_fast(cd z) {
    const double m = fmax(fabs(z.x), fabs(z.y));
    if (!(m > 0.0)) return mk(0.0, z.y);
    const int e = ((__double2hiint(m) >> 20) & 0x7ff) - 1023;          // m = f 2^e, f in [1, 2)
    const int k = e & ~1;                                              // even: sqrt(2^k) is exact
    const double sc = __hiloint2double((1023 - k) << 20, 0), bs = __hiloint2double((1023 + (k >> 1)) << 20, 0);
    const double xs = z.x * sc, ys = z.y * sc;                         // max magnitude in [1, 4)
    const double s2 = fma(xs, xs, ys * ys);
    const double az = s2 * kh_rsqrt(s2);                               // |z| scaled
    const double w = 0.5 * (fabs(xs) + az);                            // >= 0.5
    const double rw = kh_rsqrt(w);
    const double t = w * rw * bs, h = 0.5 * rw * bs;                   // sqrt(w), 1 / (2 sqrt(w)), scaled back
    if (xs >= 0.0) return mk(t, ys * h);
    return mk(fabs(ys) * h, copysign(t, ys));
}
#endif

// ------------------------------------------------------------------ CTA context
struct Cta {
    int tid, nthr;           // thread index / threads per CTA
    int bx, by;              // block indices
    unsigned char* smem;     // dynamic shared memory
#ifdef KH_HOST_EMU
    inline void sync() const {}
#else
    __device__ __forceinline__ void sync() const { __syncthreads(); }
#endif
};

#ifdef KH_HOST_EMU
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct KhProf { long long launches; };
static KhProf g_prof = {0};
#define KH_SMEM(c) ((c).smem)
template <class Args, void (*Body)(const Cta&, const Args&), int MAXT = 0, int MINB = 1>
static inline int kh_launch(dim3 grid, int /*block*/, size_t smem, kh_stream_t, const Args& a, const char* = "", double = 0.0) {
    g_prof.launches++;
    unsigned char* buf = (unsigned char*)malloc(smem + 64);
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            memset(buf, 0xFF, smem + 64);   // NaN pattern: catches reads of uninitialised shared memory
            Cta c{0, 1, (int)bx, (int)by, buf};
            Body(c, a);
        }
    free(buf);
    return 0;
}
#define KH_ATOMIC_MAX(ptr, v) (*(ptr) = std::max(*(ptr), (v)))
#define KH_ATOMIC_OR(ptr, v) (*(ptr) |= (v))
#else
extern __shared__ __align__(16) unsigned char kh_smem[];
// KH_SMEM(c): dynamic shared memory with its address space known to the compiler (LDS/STS, not generic LD/ST)
#define KH_SMEM(c) kh_smem
template <class Args, void (*Body)(const Cta&, const Args&)>
__global__ void kh_entry(const __grid_constant__ Args a) {
    Cta c{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)blockIdx.y, kh_smem};
    Body(c, a);
}
template <class Args, void (*Body)(const Cta&, const Args&), int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) kh_entry_lb(const __grid_constant__ Args a) {
    Cta c{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)blockIdx.y, kh_smem};
    Body(c, a);
}
// register budget given directly (ptxas derives a needlessly low cap from launch bounds whose thread count is not a multiple of 128)
template <class Args, void (*Body)(const Cta&, const Args&), int MAXREG>
__global__ void __maxnreg__(MAXREG) kh_entry_mr(const __grid_constant__ Args a) {
    Cta c{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)blockIdx.y, kh_smem};
    Body(c, a);
}
// launch counter + optional per-kernel CUDA-event profiler (kh_profile_begin / kh_profile_end)
#include <vector>
struct KhProfRec { const char* name; double work; cudaEvent_t e0, e1; };
struct KhProf {
    long long launches;
    bool on;
    std::vector<KhProfRec> recs;
    std::vector<cudaEvent_t> pool;
};
static KhProf g_prof = {0, false, {}, {}};
static inline cudaEvent_t kh_prof_event() {
    if (!g_prof.pool.empty()) { cudaEvent_t e = g_prof.pool.back(); g_prof.pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
template <class Args, void (*Body)(const Cta&, const Args&), int MAXT = 0, int MINB = 1>
static inline int kh_launch(dim3 grid, int block, size_t smem, kh_stream_t st, const Args& a, const char* name = "", double work = 0.0) {
    if (grid.x == 0 || grid.y == 0) return 0;
    void (*kern)(const Args);
    if constexpr (MAXT == 0) kern = kh_entry<Args, Body>;           // no launch bounds
    else if constexpr (MAXT < 0) kern = kh_entry_mr<Args, Body, MINB>;   // MAXT < 0: MINB is a register cap
    else kern = kh_entry_lb<Args, Body, MAXT, MINB>;
    if (smem > 48 * 1024) {                   // opt-in to > 48 KB dynamic smem: a per-DEVICE attribute of this instantiation
        static size_t configured[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    g_prof.launches++;
    if (g_prof.on) {
        KhProfRec r{name, work, kh_prof_event(), kh_prof_event()};
        cudaEventRecord(r.e0, st);
        kern<<<grid, block, smem, st>>>(a);
        cudaEventRecord(r.e1, st);
        g_prof.recs.push_back(r);
    } else {
        kern<<<grid, block, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}
#define KH_ATOMIC_MAX(ptr, v) atomicMax((ptr), (v))
#define KH_ATOMIC_OR(ptr, v) atomicOr((ptr), (v))
// Thread-block-cluster launch (cluster_x CTAs along x share distributed shared memory; grid.x must be a multiple of it).
template <class Args, void (*Body)(const Cta&, const Args&), int MAXT, int MINB>
static inline int kh_launch_cluster(dim3 grid, int block, size_t smem, int cluster_x, kh_stream_t st, const Args& a, const char* name = "", double work = 0.0) {
    if (grid.x == 0 || grid.y == 0) return 0;
    void (*kern)(const Args) = kh_entry_lb<Args, Body, MAXT, MINB>;
    if (smem > 48 * 1024) {
        static size_t configured[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    g_prof.launches++;
    cudaError_t e;
    if (g_prof.on) {
        KhProfRec r{name, work, kh_prof_event(), kh_prof_event()};
        cudaEventRecord(r.e0, st);
        e = cudaLaunchKernelEx(&cfg, kern, a);
        cudaEventRecord(r.e1, st);
        g_prof.recs.push_back(r);
    } else {
        e = cudaLaunchKernelEx(&cfg, kern, a);
    }
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}
#endif

// ------------------------------------------------------------------ TMA bulk copies (global -> shared, 1-D)
// Matrices that a CTA keeps resident in shared memory are staged row by row with cp.async.bulk (the TMA engine moves each row,
// completion is counted in bytes on an mbarrier) instead of per-thread loads + stores: no registers, no address arithmetic, and
// the threads are free for the padding / setup work that overlaps the copy.  Rows are 16-byte aligned multiples of 16 bytes
// (complex128), which is all the 1-D form needs -- no tensor map.
#ifndef KH_HOST_EMU
__device__ __forceinline__ unsigned kh_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kh_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(kh_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void kh_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(kh_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kh_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(kh_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(kh_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kh_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(kh_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
#endif
// Stages nrows rows of row_cd complex numbers (global, leading dimension ld_src) into shared memory (leading dimension ld_dst).
// Collective over the CTA; `bar` is 8 bytes of shared memory reserved for this call, used once (phase 0).  Ends with a barrier.
KH_DEV void kh_stage_rows(const Cta& c, cd* dst, int ld_dst, const cd* src, long long ld_src, int nrows, int row_cd, unsigned long long* bar) {
#ifdef KH_HOST_EMU
    (void)bar;
    for (int i = c.tid; i < nrows; i += c.nthr)
        for (int j = 0; j < row_cd; ++j) dst[(long long)i * ld_dst + j] = src[(long long)i * ld_src + j];
#else
    if (c.tid == 0) kh_mbar_init(bar, 1);
    __syncthreads();
    if (c.tid == 0) kh_mbar_expect_tx(bar, (unsigned)nrows * (unsigned)row_cd * 16u);
    for (int i = c.tid; i < nrows; i += c.nthr) kh_bulk_g2s(dst + (long long)i * ld_dst, src + (long long)i * ld_src, (unsigned)row_cd * 16u, bar);
    kh_mbar_wait(bar, 0);
    __syncthreads();
#endif
}

// Largest dynamic shared memory a CTA may opt in to on sm_100 (227 KB).
#define KH_SMEM_MAX (227 * 1024)

// ------------------------------------------------------------------ CTA-wide reductions
// scratch: at least 64 doubles of shared memory.  All threads get the result.
KH_DEV double cta_sum(const Cta& c, double v, double* scratch) {
#ifdef KH_HOST_EMU
    (void)c; (void)scratch;
    return v;
#else
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = c.tid >> 5, nw = (c.nthr + 31) >> 5;
    c.sync();
    if ((c.tid & 31) == 0) scratch[w] = v;
    c.sync();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += scratch[i];
    return r;
#endif
}
KH_DEV double cta_max(const Cta& c, double v, double* scratch) {
#ifdef KH_HOST_EMU
    (void)c; (void)scratch;
    return v;
#else
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = c.tid >> 5, nw = (c.nthr + 31) >> 5;
    c.sync();
    if ((c.tid & 31) == 0) scratch[w] = v;
    c.sync();
    double r = scratch[0];
    for (int i = 1; i < nw; ++i) r = fmax(r, scratch[i]);
    return r;
#endif
}
// argmax of (value, index); ties -> smallest index.  scratch: 64 doubles + 64 ints.
KH_DEV int cta_argmax(const Cta& c, double v, int idx, double* scratch) {
#ifdef KH_HOST_EMU
    (void)c; (void)scratch; (void)v;
    return idx;
#else
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    int w = c.tid >> 5, nw = (c.nthr + 31) >> 5;
    int* iscr = (int*)(scratch + 64);
    c.sync();
    if ((c.tid & 31) == 0) { scratch[w] = v; iscr[w] = idx; }
    c.sync();
    double bv = scratch[0]; int bi = iscr[0];
    for (int i = 1; i < nw; ++i)
        if (scratch[i] > bv || (scratch[i] == bv && iscr[i] < bi)) { bv = scratch[i]; bi = iscr[i]; }
    return bi;
#endif
}

// warp-level all-reduce (emulation: one virtual lane)
#ifdef KH_HOST_EMU
#define KH_WARP 1
KH_DEV double kh_warp_allsum(double v) { return v; }
#else
#define KH_WARP 32
KH_DEV double kh_warp_allsum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// batch addressing with two levels: matrix b lives at base + (b / inner) * so + (b % inner) * si
struct MatRef {
    cd* p;
    long long so, si;
    int inner, ld;
};
KH_HD MatRef mref(const void* p, long long so, int ld, int inner = 1, long long si = 0) {
    MatRef m; m.p = (cd*)p; m.so = so; m.si = si; m.inner = inner < 1 ? 1 : inner; m.ld = ld; return m;
}
KH_HD cd* mat_ptr(const MatRef& m, int b) {
    return m.p ? m.p + (long long)(b / m.inner) * m.so + (long long)(b % m.inner) * m.si : (cd*)0;
}
