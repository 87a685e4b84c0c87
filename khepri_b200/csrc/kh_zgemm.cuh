// Batched complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64).
//
//   Cout[b] = rs[i] * ( alpha * op(A[b]) * B[b] + beta * Cin[b] + diag * I ) (*|/) cs[j]
//
// tcgen05 has no FP64 kind, so warp-level DMMA is the tensor path for complex128 on sm_100a.
// A complex product is four real DMMAs per 8x8x4 step (Cr += Ar Br - Ai Bi ; Ci += Ar Bi + Ai Br).
// CTA tile 64x64, K chunks of 16 staged with cp.async (zero-filled at the ragged edges, so
// matrices need no padding in HBM), double buffered.  Eight warps; warp w owns rows 8w..8w+7 of
// the tile and all eight 8-column MMA tiles.  Shared-memory strides (20 / 66 complex) make the
// A- and B-fragment loads bank-conflict free.
#pragma once
#include "kh_common.cuh"

struct zgemm_args {
    int M, N, K;
    int transA;                 // 1: A is stored K x M (row-major) and used transposed
    MatRef A, B, Cin, Cout;     // Cin.p may be null
    double alpha, beta, diag;
    const cd* rowscale;         // optional, length M per batch group
    long long rs_stride; int rs_group;
    const cd* colscale;         // optional, length N per batch group
    long long cs_stride; int cs_group;
    int cs_divide;              // 1: divide by colscale instead of multiplying
};

#define ZG_BM 64
#define ZG_BN 64
#define ZG_BK 16
#define ZG_LDA 20
#define ZG_LDB 66
#define ZG_THREADS 256
#define ZG_SMEM (2 * (ZG_BM * ZG_LDA + ZG_BK * ZG_LDB) * (int)sizeof(cd))

KH_DEV cd zgemm_epilogue(const zgemm_args& a, const cd* cin, const cd* rs, const cd* cs, int row, int col, cd acc) {
    cd v = a.alpha * acc;
    if (cin) v = v + a.beta * cin[(long long)row * a.Cin.ld + col];
    if (a.diag != 0.0 && row == col) v.x += a.diag;
    if (rs) v = rs[row] * v;
    if (cs) v = a.cs_divide ? v / cs[col] : v * cs[col];
    return v;
}

#ifndef KH_HOST_EMU
__device__ __forceinline__ void kh_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void kh_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void kh_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void kh_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

KH_DEV void zgemm_body(const Cta& c, const zgemm_args& a) {
    const int b = c.bx;
    const int tiles_n = (a.N + ZG_BN - 1) / ZG_BN;
    const int m0 = (c.by / tiles_n) * ZG_BM, n0 = (c.by % tiles_n) * ZG_BN;
    const cd* A = mat_ptr(a.A, b);
    const cd* B = mat_ptr(a.B, b);
    const cd* Cin = mat_ptr(a.Cin, b);
    cd* Cout = mat_ptr(a.Cout, b);
    const cd* rs = a.rowscale ? a.rowscale + (long long)(b / a.rs_group) * a.rs_stride : (const cd*)0;
    const cd* cs = a.colscale ? a.colscale + (long long)(b / a.cs_group) * a.cs_stride : (const cd*)0;
#ifdef KH_HOST_EMU
    for (int i = m0; i < m0 + ZG_BM && i < a.M; ++i)
        for (int j = n0; j < n0 + ZG_BN && j < a.N; ++j) {
            cd acc = mk(0, 0);
            for (int k = 0; k < a.K; ++k) {
                cd av = a.transA ? A[(long long)k * a.A.ld + i] : A[(long long)i * a.A.ld + k];
                cfma(acc, av, B[(long long)k * a.B.ld + j]);
            }
            Cout[(long long)i * a.Cout.ld + j] = zgemm_epilogue(a, Cin, rs, cs, i, j, acc);
        }
#else
    cd* As = (cd*)c.smem;                               // [2][BM][LDA]
    cd* Bs = As + 2 * ZG_BM * ZG_LDA;                   // [2][BK][LDB]
    const int warp = c.tid >> 5, lane = c.tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    double cr[8][2], ci[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0; }
    const int nk = (a.K + ZG_BK - 1) / ZG_BK;
    const int nt = min(8, (a.N - n0 + 7) >> 3);
    const bool warp_active = (m0 + warp * 8) < a.M;

    auto stage = [&](int buf, int kc) {
        const int k0 = kc * ZG_BK;
        cd* as = As + buf * ZG_BM * ZG_LDA;
        cd* bs = Bs + buf * ZG_BK * ZG_LDB;
        if (!a.transA) {
#pragma unroll
            for (int e = c.tid; e < ZG_BM * ZG_BK; e += ZG_THREADS) {
                int m = e >> 4, k = e & 15;
                bool ok = (m0 + m) < a.M && (k0 + k) < a.K;
                kh_cp_async16(as + m * ZG_LDA + k, ok ? A + (long long)(m0 + m) * a.A.ld + k0 + k : A, ok);
            }
        } else {
#pragma unroll
            for (int e = c.tid; e < ZG_BM * ZG_BK; e += ZG_THREADS) {
                int k = e >> 6, m = e & 63;
                bool ok = (m0 + m) < a.M && (k0 + k) < a.K;
                kh_cp_async16(as + m * ZG_LDA + k, ok ? A + (long long)(k0 + k) * a.A.ld + m0 + m : A, ok);
            }
        }
#pragma unroll
        for (int e = c.tid; e < ZG_BK * ZG_BN; e += ZG_THREADS) {
            int k = e >> 6, n = e & 63;
            bool ok = (k0 + k) < a.K && (n0 + n) < a.N;
            kh_cp_async16(bs + k * ZG_LDB + n, ok ? B + (long long)(k0 + k) * a.B.ld + n0 + n : B, ok);
        }
        kh_cp_async_commit();
    };

    stage(0, 0);
    for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) { stage(buf ^ 1, kc + 1); kh_cp_async_wait<1>(); }
        else kh_cp_async_wait<0>();
        __syncthreads();
        if (warp_active) {
            const cd* as = As + buf * ZG_BM * ZG_LDA + (warp * 8 + lr) * ZG_LDA + lk;
            const cd* bs = Bs + buf * ZG_BK * ZG_LDB + lk * ZG_LDB + lr;
#pragma unroll
            for (int kk = 0; kk < ZG_BK / 4; ++kk) {
                cd av = as[kk * 4];
                double nai = -av.y;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (t < nt) {
                        cd bv = bs[kk * 4 * ZG_LDB + t * 8];
                        kh_dmma(cr[t][0], cr[t][1], av.x, bv.x);
                        kh_dmma(cr[t][0], cr[t][1], nai, bv.y);
                        kh_dmma(ci[t][0], ci[t][1], av.x, bv.y);
                        kh_dmma(ci[t][0], ci[t][1], av.y, bv.x);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (warp_active) {
        const int row = m0 + warp * 8 + lr;
        if (row < a.M) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (t < nt) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int col = n0 + t * 8 + 2 * lk + h;
                        if (col < a.N)
                            Cout[(long long)row * a.Cout.ld + col] = zgemm_epilogue(a, Cin, rs, cs, row, col, mk(cr[t][h], ci[t][h]));
                    }
                }
            }
        }
    }
#endif
}

static inline int zgemm_launch(kh_stream_t st, int batch, const zgemm_args& a) {
    if (batch <= 0 || a.M <= 0 || a.N <= 0) return 0;
    int tiles = ((a.M + ZG_BM - 1) / ZG_BM) * ((a.N + ZG_BN - 1) / ZG_BN);
    return kh_launch<zgemm_args, zgemm_body>(dim3(batch, tiles), ZG_THREADS, ZG_SMEM, st, a, "zgemm", 8.0 * a.M * a.N * a.K * batch);
}

// convenience builder: plain C = alpha*A*B (+ beta*Cin) on [batch, n, n] row-major stacks
static inline zgemm_args zgemm_make(int M, int N, int K, MatRef A, MatRef B, MatRef Cout, double alpha = 1.0) {
    zgemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.A = A; g.B = B; g.Cout = Cout;
    g.Cin = mref(nullptr, 0, 0);
    g.alpha = alpha; g.beta = 0.0; g.diag = 0.0;
    g.rs_group = g.cs_group = 1;
    return g;
}
