// Batched complex128 matrix inverse: in-place Gauss-Jordan with partial (row) pivoting,
// one CTA per matrix.  The matrix lives in shared memory when it fits (n <= ~118), otherwise the
// same code runs directly on the output matrix in HBM/L2 (slow path for the large extended-RCWA
// bases; replaced by a blocked variant later).
//
// Every dense "solve(A, B)" of the reference (numpy.linalg.solve, alternative.py:24-27, 182-193)
// becomes inverse + DMMA GEMM here; pivot choice follows LAPACK's izamax (|re|+|im|).
#pragma once
#include "kh_common.cuh"
#include "kh_zgemm.cuh"

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
    constexpr int L = (TR == 2) ? 10 : ((TR == 3) ? 15 : 20);             // lcm(TR, 5)
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
__device__ __forceinline__ void zinv_reg_t3_body(const Cta& c, const zinv_args& a) { zinv_reg_body<3>(c, a); }
__device__ __forceinline__ void zinv_reg_t2x_body(const Cta& c, const zinv_args& a) { zinv_reg_body<2>(c, a); }
#endif

// ---------------------------------------------------------------------------------------------
// Tiled variant for matrices that do not fit in shared memory (n > ~118: the 9x9 ... 15x15 harmonic
// bases and the extended-RCWA supercells): BLOCKED in-place Gauss-Jordan.  Per block column of nb
// pivots:
//   panel kernel (one CTA per matrix): the n x nb block column is staged in shared memory and
//     eliminated there with partial pivoting (pivot rows from the diagonal block downwards), which
//     turns it into the block column  P' = [-A01 A11^-1; A11^-1; -A21 A11^-1]  of the Gauss-Jordan
//     transformation.  The row interchanges are then applied to the rest of the matrix, the nb pivot
//     rows R (all other columns) are moved to a side buffer and zeroed in place;
//   two DMMA GEMMs:  A[:, left] += P' R[:, left],  A[:, right] += P' R[:, right]   (M = n, K = nb).
// All 8 n^3 flops of the inversion run on the tensor pipe; HBM traffic is one read+write of the
// matrix per block column.  A final kernel undoes the row interchanges as one column gather.
struct zinvb_args {
    int n, k0, nb;
    MatRef A;                         // in-place matrix
    cd* R; long long r_stride;        // [batch][nb][n] pivot rows of the current step
    int* piv; long long piv_stride;   // [batch][n] pivot rows (absolute)
    int* info;
    int lds;
};

KH_DEV void zinvb_panel_body(const Cta& c, const zinvb_args& a) {
    const int n = a.n, b = c.bx, k0 = a.k0, lds = a.lds;
    const int nb = (a.nb < n - k0) ? a.nb : n - k0;
    cd* A = mat_ptr(a.A, b);
    const int ld = a.A.ld;
    cd* R = a.R + (long long)b * a.r_stride;
    int* pivg = a.piv + (long long)b * a.piv_stride;
    // shared: [colk n][rowk 32][scratch 128 dbl][pivs 32 int][panel n x lds]
    cd* colk = (cd*)KH_SMEM(c);
    cd* rowk = colk + n;
    double* scratch = (double*)(rowk + 32);
    int* pivs = (int*)(scratch + 128);
    cd* Ps = (cd*)(KH_SMEM(c) + (((n * 16 + 32 * 16 + 128 * 8 + 32 * 4) + 15) & ~15));
    for (int e = c.tid; e < n * nb; e += c.nthr) { int i = e / nb, j = e - i * nb; Ps[i * lds + j] = A[(long long)i * ld + k0 + j]; }
    c.sync();
    int bad = 0;
    for (int s = 0; s < nb; ++s) {
        const int k = k0 + s;
        double best = -1.0; int bi = k;
        for (int i = k + c.tid; i < n; i += c.nthr) {
            double v = cabs1(Ps[i * lds + s]);
            if (v > best) { best = v; bi = i; }
        }
        const int p = cta_argmax(c, best, bi, scratch);
        if (c.tid == 0) pivs[s] = p;
        if (p != k)
            for (int j = c.tid; j < nb; j += c.nthr) { cd t = Ps[k * lds + j]; Ps[k * lds + j] = Ps[p * lds + j]; Ps[p * lds + j] = t; }
        c.sync();
        const cd pv = Ps[k * lds + s];
        if (pv.x == 0.0 && pv.y == 0.0 && !bad) bad = k + 1;
        const cd d = crecip(pv);
        for (int i = c.tid; i < n; i += c.nthr) colk[i] = Ps[i * lds + s];
        for (int j = c.tid; j < nb; j += c.nthr) rowk[j] = (j == s) ? d : Ps[k * lds + j] * d;
        c.sync();
        for (int e = c.tid; e < n * nb; e += c.nthr) {
            const int i = e / nb, j = e - i * nb;
            cd v;
            if (i == k) v = rowk[j];
            else if (j == s) v = -(colk[i] * d);
            else { v = Ps[i * lds + j]; cfms(v, colk[i], rowk[j]); }
            Ps[i * lds + j] = v;
        }
        c.sync();
    }
    // row interchanges on the columns outside the panel (sequential: later swaps may touch the same rows)
    const int nout = n - nb;
    for (int s = 0; s < nb; ++s) {
        const int k = k0 + s, p = pivs[s];
        if (p != k) {
            for (int t = c.tid; t < nout; t += c.nthr) {
                const int j = t < k0 ? t : t + nb;
                cd x = A[(long long)k * ld + j]; A[(long long)k * ld + j] = A[(long long)p * ld + j]; A[(long long)p * ld + j] = x;
            }
            c.sync();
        }
    }
    // panel out; pivot rows to the side buffer, zero in place
    for (int e = c.tid; e < n * nb; e += c.nthr) { int i = e / nb, j = e - i * nb; A[(long long)i * ld + k0 + j] = Ps[i * lds + j]; }
    for (int e = c.tid; e < nb * nout; e += c.nthr) {
        const int s = e / nout, t = e - s * nout, j = t < k0 ? t : t + nb;
        R[(long long)s * n + j] = A[(long long)(k0 + s) * ld + j];
        A[(long long)(k0 + s) * ld + j] = mk(0.0, 0.0);
    }
    for (int s = c.tid; s < nb; s += c.nthr) pivg[k0 + s] = pivs[s];
    if (a.info && c.tid == 0) { if (k0 == 0) a.info[b] = bad; else if (bad && a.info[b] == 0) a.info[b] = bad; }
}

// undo the row interchanges: out[:, j] = in[:, src[j]] with src = the column swaps (k <-> piv[k]) applied for k = n-1 .. 0
KH_DEV void zinvb_unpermute_body(const Cta& c, const zinvb_args& a) {
    const int n = a.n, b = c.bx;
    cd* A = mat_ptr(a.A, b);
    const int ld = a.A.ld;
    const int* pivg = a.piv + (long long)b * a.piv_stride;
    int* src = (int*)KH_SMEM(c);
    const int nw = (c.nthr + KH_WARP - 1) / KH_WARP, warp = c.tid / KH_WARP, lane = c.tid % KH_WARP;
    cd* rows = (cd*)(KH_SMEM(c) + (((n * 4) + 15) & ~15));           // one row buffer per warp
    for (int j = c.tid; j < n; j += c.nthr) src[j] = j;
    c.sync();
    if (c.tid == 0)
        for (int k = n - 1; k >= 0; --k) { const int p = pivg[k]; if (p != k) { int t = src[k]; src[k] = src[p]; src[p] = t; } }
    c.sync();
    cd* buf = rows + (long long)warp * n;
    for (int i = warp; i < n; i += nw) {
        for (int j = lane; j < n; j += KH_WARP) buf[j] = A[(long long)i * ld + j];
#ifndef KH_HOST_EMU
        __syncwarp();
#endif
        for (int j = lane; j < n; j += KH_WARP) A[(long long)i * ld + j] = buf[src[j]];
#ifndef KH_HOST_EMU
        __syncwarp();
#endif
    }
}

struct zcopym_args { int n; MatRef src, dst; };
KH_DEV void zcopym_body(const Cta& c, const zcopym_args& a) {
    const cd* s = mat_ptr(a.src, c.bx); cd* d = mat_ptr(a.dst, c.bx);
    for (int e = c.tid; e < a.n * a.n; e += c.nthr) { int i = e / a.n, j = e - i * a.n; d[(long long)i * a.dst.ld + j] = s[(long long)i * a.src.ld + j]; }
}

static inline int zinvb_nb(int n) {
    int nb = 32;
    while (nb > 4 && (size_t)n * (nb + 1) * sizeof(cd) + (size_t)n * 16 + 4096 > (size_t)200 * 1024) nb >>= 1;
    return nb;
}
// work space of the blocked variant, in complex elements per matrix (pivot rows + pivot indices)
static inline long long zinv_work_cd(int n) { return (long long)zinvb_nb(n) * n + (n + 3) / 4 + 4; }
#ifndef KH_ZINV_BLOCKED_MIN
#define KH_ZINV_BLOCKED_MIN 101
#endif

static inline int zinv_blocked_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info, cd* work) {
    int e;
    if (A.p != Ainv.p) {
        zcopym_args cp{n, A, Ainv};
        if ((e = kh_launch<zcopym_args, zcopym_body>(dim3(batch), 256, 0, st, cp, "zinv", 0.0))) return e;
    }
    const int nb = zinvb_nb(n);
    const long long wstride = zinv_work_cd(n);
    zinvb_args a;
    a.n = n; a.nb = nb; a.A = Ainv; a.R = work; a.r_stride = wstride; a.info = info; a.lds = nb + 1;
    a.piv = (int*)(work + (long long)nb * n); a.piv_stride = wstride * 4;
    const size_t sm = (size_t)n * 16 + 32 * 16 + 128 * 8 + 32 * 4 + 16 + (size_t)n * a.lds * sizeof(cd);
    for (int k0 = 0; k0 < n; k0 += nb) {
        a.k0 = k0;
        const int nbk = nb < n - k0 ? nb : n - k0;
        if ((e = kh_launch<zinvb_args, zinvb_panel_body>(dim3(batch), 512, sm, st, a, "zinv", 0.0))) return e;
        MatRef Pm = Ainv; Pm.p = Ainv.p + k0;
        MatRef Rm = mref(work, wstride, n);
        if (k0 > 0) {
            zgemm_args g = zgemm_make(n, k0, nbk, Pm, Rm, Ainv);
            g.Cin = Ainv; g.beta = 1.0;
            if ((e = zgemm_launch(st, batch, g))) return e;
        }
        if (k0 + nbk < n) {
            MatRef Rr = Rm; Rr.p = work + k0 + nbk;
            MatRef Cr = Ainv; Cr.p = Ainv.p + k0 + nbk;
            zgemm_args g = zgemm_make(n, n - k0 - nbk, nbk, Pm, Rr, Cr);
            g.Cin = Cr; g.beta = 1.0;
            if ((e = zgemm_launch(st, batch, g))) return e;
        }
    }
    const int uthr = 256;
    const size_t usm = (size_t)((n * 4 + 15) & ~15) + (size_t)(uthr / 32) * n * sizeof(cd) + 16;
    return kh_launch<zinvb_args, zinvb_unpermute_body>(dim3(batch), uthr, usm, st, a, "zinv", 0.0);
}

static inline size_t zinv_smem_bytes(int n, int ld_s, int use_smem) {
    size_t s = (size_t)2 * n * sizeof(cd) + 128 * sizeof(double) + (size_t)n * sizeof(int) + 16;
    if (use_smem) s += (size_t)n * ld_s * sizeof(cd);
    return s;
}

static inline int zinv_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info, cd* work = nullptr, long long work_cd = 0) {
    if (batch <= 0 || n <= 0) return 0;
    if (n >= KH_ZINV_BLOCKED_MIN && work && work_cd >= (long long)batch * zinv_work_cd(n)) return zinv_blocked_launch(st, batch, n, A, Ainv, info, work);
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
#ifndef KH_ZINV_VARIANT
#define KH_ZINV_VARIANT 0
#endif
        const int tiles3 = txn * ((n + 2) / 3);
        if (KH_ZINV_VARIANT == 1 && tiles3 <= 704) return kh_launch<zinv_args, zinv_reg_t3_body, 704, 1>(dim3(batch), ((tiles3 + 31) / 32) * 32, sm, st, a, "zinv", work);
        if (KH_ZINV_VARIANT == 2 && tiles2 <= 1024) return kh_launch<zinv_args, zinv_reg_t2x_body, 1024, 1>(dim3(batch), ((tiles2 + 31) / 32) * 32, sm, st, a, "zinv", work);
        return kh_launch<zinv_args, zinv_reg_large_body, 512, 1>(dim3(batch), ((tiles4 + 31) / 32) * 32, sm, st, a, "zinv", work);
    }
#endif
    a.ld_s = n | 1;
    a.use_smem = zinv_smem_bytes(n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int threads = n <= 64 ? 256 : 512;
    return kh_launch<zinv_args, zinv_body>(dim3(batch), threads, zinv_smem_bytes(n, a.ld_s, a.use_smem), st, a, "zinv", 8.0 * n * n * n * batch);
}
