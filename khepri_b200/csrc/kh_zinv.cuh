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
    cd* colk = (cd*)c.smem;
    cd* rowk = colk + n;
    double* scratch = (double*)(rowk + n);
    int* piv = (int*)(scratch + 128);
    cd* W;
    int ld;
    if (a.use_smem) {
        W = (cd*)(((uintptr_t)(piv + n) + 15) & ~(uintptr_t)15);
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

static inline size_t zinv_smem_bytes(int n, int ld_s, int use_smem) {
    size_t s = (size_t)2 * n * sizeof(cd) + 128 * sizeof(double) + (size_t)n * sizeof(int) + 16;
    if (use_smem) s += (size_t)n * ld_s * sizeof(cd);
    return s;
}

static inline int zinv_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info) {
    if (batch <= 0 || n <= 0) return 0;
    zinv_args a;
    a.n = n; a.A = A; a.Ainv = Ainv; a.info = info;
    a.ld_s = n | 1;
    a.use_smem = zinv_smem_bytes(n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int threads = n <= 64 ? 256 : 512;
    return kh_launch<zinv_args, zinv_body>(dim3(batch), threads, zinv_smem_bytes(n, a.ld_s, a.use_smem), st, a, "zinv", 8.0 * n * n * n * batch);
}
