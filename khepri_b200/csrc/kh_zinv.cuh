// Batched complex128 matrix inverse: in-place Gauss-Jordan with partial (row) pivoting,
// one CTA per matrix.  The matrix lives in shared memory when it fits (n <= ~118), otherwise the
// same code runs directly on the output matrix in HBM/L2 (slow path for the large extended-RCWA
// bases; replaced by a blocked variant later).
//
// Every dense "solve(A, B)" of the reference (numpy.linalg.solve, alternative.py:24-27, 182-193)
// becomes inverse + DMMA GEMM here; pivot choice follows LAPACK's izamax (|re|+|im|).
#pragma once
#include "kh_common.cuh"

struct zinv_args {
    int n;
    MatRef A, Ainv;      // Ainv may alias A
    int* info;           // per-matrix status (0 ok, k+1 = zero pivot at step k); may be null
    int use_smem;
    int ld_s;            // shared-memory leading dimension (odd)
};

KH_DEV void zinv_body(const Cta& c, const zinv_args& a) {
    const int n = a.n, b = c.bx;
    const cd* A = mat_ptr(a.A, b);
    cd* Out = mat_ptr(a.Ainv, b);
    // shared layout: [colk n][rowk n][scratch 128 dbl][piv n ints][matrix]
    cd* colk = (cd*)KH_SMEM(c);
    cd* rowk = colk + n;
    double* scratch = (double*)(rowk + n);
    int* piv = (int*)(scratch + 128);
    cd* W;
    int ld;
    if (a.use_smem) {
        W = (cd*)(KH_SMEM(c) + (((2 * n * 16 + 128 * 8 + n * 4) + 15) & ~15));   // offset arithmetic keeps the shared address space
        ld = a.ld_s;
    } else {
        W = Out;
        ld = a.Ainv.ld;
    }
    if (a.use_smem || A != Out) {
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            W[(long long)i * ld + j] = A[(long long)i * a.A.ld + j];
        }
    }
    c.sync();
    int bad = 0;
    for (int k = 0; k < n; ++k) {
        // pivot search in column k, rows k..n-1
        double best = -1.0; int bi = k;
        for (int i = k + c.tid; i < n; i += c.nthr) {
            double v = cabs1(W[(long long)i * ld + k]);
            if (v > best) { best = v; bi = i; }
        }
        int p = cta_argmax(c, best, bi, scratch);
        if (c.tid == 0) piv[k] = p;
        // swap rows k and p
        if (p != k) {
            for (int j = c.tid; j < n; j += c.nthr) {
                cd t = W[(long long)k * ld + j];
                W[(long long)k * ld + j] = W[(long long)p * ld + j];
                W[(long long)p * ld + j] = t;
            }
        }
        c.sync();
        cd pv = W[(long long)k * ld + k];
        if (pv.x == 0.0 && pv.y == 0.0 && !bad) bad = k + 1;
        cd d = crecip(pv);
        // stash column k and the scaled pivot row
        for (int i = c.tid; i < n; i += c.nthr) {
            colk[i] = W[(long long)i * ld + k];
            rowk[i] = (i == k) ? d : W[(long long)k * ld + i] * d;
        }
        c.sync();
        // rank-1 update of everything except pivot row / column, which are set directly
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            cd v;
            if (i == k) v = rowk[j];
            else if (j == k) v = -(colk[i] * d);
            else { v = W[(long long)i * ld + j]; cfms(v, colk[i], rowk[j]); }
            W[(long long)i * ld + j] = v;
        }
        c.sync();
    }
    // undo the row interchanges as column interchanges, in reverse order
    for (int k = n - 1; k >= 0; --k) {
        int p = piv[k];
        if (p != k) {
            for (int i = c.tid; i < n; i += c.nthr) {
                cd t = W[(long long)i * ld + k];
                W[(long long)i * ld + k] = W[(long long)i * ld + p];
                W[(long long)i * ld + p] = t;
            }
        }
        c.sync();
    }
    if (a.use_smem) {
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            Out[(long long)i * a.Ainv.ld + j] = W[(long long)i * ld + j];
        }
    }
    if (a.info && c.tid == 0) a.info[b] = bad;
}


#ifndef KH_HOST_EMU
// ---------------------------------------------------------------------------------------------
// Register-resident variant for n <= 100 (the 5x5 and 7x7 harmonic bases): the whole matrix lives
// in the register file, each thread owning a TR x 5 tile (the matrix is padded with identity rows /
// columns up to the tile grid, so the elimination loop has no bounds checks).  Per step only the
// pivot column and the two rows of the interchange travel through shared memory (double buffered:
// two barriers per step); the rank-1 update is pure register DFMA work.  Because tiles are aligned,
// the pivot column / row of step k sit in register slot (k mod 5, k mod TR) of their owners, so the
// step loop is unrolled by lcm(TR, 5) and every register index is a compile-time constant.
#define ZIR_TC 5
template <int TR, int KQ, int KP>
__device__ __forceinline__ void zir_step(int k, int n, int NP, int tid, int lane, bool live, int i0, int j0, int tx, int ty,
                                         cd (&r)[TR][ZIR_TC], cd* colk, cd* rowK, cd* rowP, int* piv, int& bad) {
    cd* ck = colk + (k & 1) * NP; cd* rK = rowK + (k & 1) * NP; cd* rP = rowP + (k & 1) * NP;
    const bool own_col = live && (j0 + KQ == k), own_row = live && (i0 + KP == k);
    // A: publish column k and row k
    if (own_col) {
#pragma unroll
        for (int p = 0; p < TR; ++p) ck[i0 + p] = r[p][KQ];
    }
    if (own_row) {
#pragma unroll
        for (int q = 0; q < ZIR_TC; ++q) rK[j0 + q] = r[KP][q];
    }
    __syncthreads();
    // B: every warp finds the pivot row (izamax over rows k..n-1, ties -> smallest index)
    double best = -1.0; int pr = k;
    for (int i = k + lane; i < n; i += 32) { const double v = cabs1(ck[i]); if (v > best) { best = v; pr = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, pr, o);
        if (ov > best || (ov == best && oi < pr)) { best = ov; pr = oi; }
    }
    const int pp = pr - i0;
    const bool own_prow = live && pr != k && (unsigned)pp < (unsigned)TR;
    if (__any_sync(0xffffffffu, own_prow)) {                 // rare per warp: a real branch
        if (own_prow) {
#pragma unroll
            for (int q = 0; q < ZIR_TC; ++q) {
                cd v = r[0][q];
#pragma unroll
                for (int p2 = 1; p2 < TR; ++p2) if (pp == p2) v = r[p2][q];
                rP[j0 + q] = v;
            }
        }
    }
    if (tid == 0) piv[k] = pr;
    __syncthreads();
    // C: interchange + eliminate.  Row k becomes the scaled pivot row; row pr receives the old row k.
    const cd* prow = (pr == k) ? rK : rP;
    const cd pv = ck[pr];
    if (pv.x == 0.0 && pv.y == 0.0 && !bad) bad = k + 1;
    const cd d = crecip(pv);
    cd f[TR];
#pragma unroll
    for (int p = 0; p < TR; ++p) f[p] = ck[i0 + p];
    if (__any_sync(0xffffffffu, own_prow)) {
        if (own_prow) {
#pragma unroll
            for (int p = 0; p < TR; ++p) if (p == pp) {
                f[p] = ck[k];
#pragma unroll
                for (int q = 0; q < ZIR_TC; ++q) r[p][q] = rK[j0 + q];
            }
        }
    }
    if (own_col) {
#pragma unroll
        for (int p = 0; p < TR; ++p) r[p][KQ] = mk(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < ZIR_TC; ++q) {
        cd pj = prow[j0 + q] * d;
        if (q == KQ && own_col) pj = d;
#pragma unroll
        for (int p = 0; p < TR; ++p) cfms(r[p][q], f[p], pj);
        if (own_row) r[KP][q] = pj;
    }
    // (no barrier here: the next step writes the other buffer set)
}

template <int TR, int U, int L>
struct zir_unroll {
    static __device__ __forceinline__ void run(int kb, int n, int NP, int tid, int lane, bool live, int i0, int j0, int tx, int ty,
                                               cd (&r)[TR][ZIR_TC], cd* colk, cd* rowK, cd* rowP, int* piv, int& bad) {
        if (kb + U < n) {
            zir_step<TR, U % ZIR_TC, U % TR>(kb + U, n, NP, tid, lane, live, i0, j0, tx, ty, r, colk, rowK, rowP, piv, bad);
            zir_unroll<TR, U + 1, L>::run(kb, n, NP, tid, lane, live, i0, j0, tx, ty, r, colk, rowK, rowP, piv, bad);
        }
    }
};
template <int TR, int L>
struct zir_unroll<TR, L, L> {
    static __device__ __forceinline__ void run(int, int, int, int, int, bool, int, int, int, int, cd (&)[TR][ZIR_TC], cd*, cd*, cd*, int*, int&) {}
};

template <int TR>
__device__ __forceinline__ void zinv_reg_body(const Cta& c, const zinv_args& a) {
    constexpr int L = (TR == 2) ? 10 : 20;             // lcm(TR, 5)
    const int n = a.n, b = c.bx, tid = c.tid, lane = tid & 31;
    const cd* A = mat_ptr(a.A, b);
    cd* Out = mat_ptr(a.Ainv, b);
    const int TXN = (n + ZIR_TC - 1) / ZIR_TC, TYN = (n + TR - 1) / TR;
    const int NP = max(TXN * ZIR_TC, TYN * TR);
    const bool live = tid < TXN * TYN;
    const int ty = live ? tid / TXN : 0, tx = live ? tid - ty * TXN : 0;
    const int i0 = ty * TR, j0 = tx * ZIR_TC;
    cd* colk = (cd*)KH_SMEM(c);               // [2][NP]
    cd* rowK = colk + 2 * NP;                 // [2][NP]
    cd* rowP = rowK + 2 * NP;                 // [2][NP]
    int* piv = (int*)(rowP + 2 * NP);         // [n]
    int* dest = piv + n;                      // [n]
    cd r[TR][ZIR_TC];
#pragma unroll
    for (int p = 0; p < TR; ++p)
#pragma unroll
        for (int q = 0; q < ZIR_TC; ++q) {
            const int i = i0 + p, j = j0 + q;
            r[p][q] = (live && i < n && j < n) ? A[(long long)i * a.A.ld + j] : mk((i == j) ? 1.0 : 0.0, 0.0);
        }
    int bad = 0;
    for (int kb = 0; kb < n; kb += L)
        zir_unroll<TR, 0, L>::run(kb, n, NP, tid, lane, live, i0, j0, tx, ty, r, colk, rowK, rowP, piv, bad);
    __syncthreads();
    // undo the row interchanges as column interchanges (reverse order) -> destination column of every stored column
    if (tid == 0) {
        for (int j = 0; j < n; ++j) dest[j] = j;                  // dest[pos] = stored column sitting at pos
        for (int k = n - 1; k >= 0; --k) { const int p = piv[k]; const int t = dest[k]; dest[k] = dest[p]; dest[p] = t; }
        for (int pos = 0; pos < n; ++pos) piv[dest[pos]] = pos;   // invert (piv is free now)
        for (int j = 0; j < n; ++j) dest[j] = piv[j];
    }
    __syncthreads();
    if (live) {
#pragma unroll
        for (int p = 0; p < TR; ++p)
#pragma unroll
            for (int q = 0; q < ZIR_TC; ++q) {
                const int i = i0 + p, j = j0 + q;
                if (i < n && j < n) Out[(long long)i * a.Ainv.ld + dest[j]] = r[p][q];
            }
    }
    if (a.info && tid == 0) a.info[b] = bad;
}
__device__ __forceinline__ void zinv_reg_small_body(const Cta& c, const zinv_args& a) { zinv_reg_body<2>(c, a); }
__device__ __forceinline__ void zinv_reg_mid_body(const Cta& c, const zinv_args& a) { zinv_reg_body<2>(c, a); }
__device__ __forceinline__ void zinv_reg_large_body(const Cta& c, const zinv_args& a) { zinv_reg_body<4>(c, a); }
#endif

static inline size_t zinv_smem_bytes(int n, int ld_s, int use_smem) {
    size_t s = (size_t)2 * n * sizeof(cd) + 128 * sizeof(double) + (size_t)n * sizeof(int) + 16;
    if (use_smem) s += (size_t)n * ld_s * sizeof(cd);
    return s;
}

static inline int zinv_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info) {
    if (batch <= 0 || n <= 0) return 0;
    zinv_args a;
    a.n = n; a.A = A; a.Ainv = Ainv; a.info = info;
#ifndef KH_HOST_EMU
    if (n <= 100) {
        a.use_smem = 0; a.ld_s = 0;
        const int txn = (n + ZIR_TC - 1) / ZIR_TC;
        const int tiles2 = txn * ((n + 1) / 2), tiles4 = txn * ((n + 3) / 4);
        const size_t sm = (size_t)6 * (n + 8) * sizeof(cd) + (size_t)2 * n * sizeof(int) + 16;
        const double work = 8.0 * n * n * n * batch;
        if (tiles2 <= 256) return kh_launch<zinv_args, zinv_reg_small_body, 256, 2>(dim3(batch), ((tiles2 + 31) / 32) * 32, sm, st, a, "zinv", work);
        if (tiles2 <= 512) return kh_launch<zinv_args, zinv_reg_mid_body, 512, 1>(dim3(batch), ((tiles2 + 31) / 32) * 32, sm, st, a, "zinv", work);
        return kh_launch<zinv_args, zinv_reg_large_body, 512, 1>(dim3(batch), ((tiles4 + 31) / 32) * 32, sm, st, a, "zinv", work);
    }
#endif
    a.ld_s = n | 1;
    a.use_smem = zinv_smem_bytes(n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int threads = n <= 64 ? 256 : 512;
    return kh_launch<zinv_args, zinv_body>(dim3(batch), threads, zinv_smem_bytes(n, a.ld_s, a.use_smem), st, a, "zinv", 8.0 * n * n * n * batch);
}
