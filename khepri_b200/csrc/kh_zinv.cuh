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

// status bookkeeping shared by the inverse kernels (see zinv_args::info_mode)
KH_DEV void zinv_note(int* info, int mode, int b, int bad) {
    if (!info) return;
    if (mode == 0) info[b] = bad;
    else if (bad) KH_ATOMIC_OR(&info[b / mode], 2);
}

struct zinv_args {
    int n;
    MatRef A, Ainv;      // Ainv may alias A
    int* info;           // per-matrix status (0 ok, k+1 = zero pivot at step k); may be null
    int info_mode = 0;   // 0: info[b] = status.  m > 0: accumulate, info[b / m] |= 2 on a zero pivot (pipeline use: the solve's info word)
    int use_smem;
    int ld_s;            // shared-memory leading dimension (odd)
    cd* gwork = nullptr; // zinv_l2: [batch][np][np] working copies in global memory
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
    if (c.tid == 0) zinv_note(a.info, a.info_mode, b, bad);
}



#ifndef KH_HOST_EMU
// ---------------------------------------------------------------------------------------------
// Shared-memory resident BLOCKED Gauss-Jordan for n <= 104 (the 5x5 and 7x7 harmonic bases): one CTA per
// matrix, the matrix (padded with identity to a multiple of 8) stays in shared memory for the whole
// inversion.  Per block column of NB = 16 pivots:
//   panel  : warps 0-3, thread <-> row, the row's NB panel entries in registers; per pivot a warp-shuffle
//            argmax + one 4-entry exchange, the pivot row broadcast through shared memory, two 128-thread
//            named barriers; the eliminated panel is the block column P' of the Gauss-Jordan transformation;
//   rows   : thread <-> column applies the panel's interchanges to the other columns and moves the NB pivot
//            rows R to a side buffer (zeroed in place);
//   update : all 16 warps, A[:, other] += P' R on 8x8 DMMA tiles, C read from / written to shared memory.
// All 8 n^3 flops of the rank updates run as DMMA (same pipe throughput as DFMA on B200, but 1/8 of the
// issue slots and no per-pivot CTA-wide barriers), which is what the register-resident variant was short of.
// 1/z without divisions (FP64 division is a ~25-deep dependent chain at ~20 cycles per op on B200): exact power-of-two
// scaling, |z|^2, hardware reciprocal seed (20 bits) + two Newton steps.  z = 0 gives NaN/Inf like the division would.
__device__ __forceinline__ cd kh_crecip_fast(cd z) {
    const double m = fmax(fabs(z.x), fabs(z.y));
    const int e = (__double2hiint(m) >> 20) & 0x7ff;
    const double sc = __hiloint2double((2046 - e) << 20, 0);          // 2^(1023 - e): scaled max magnitude in [1, 2)
    const double xs = z.x * sc, ys = z.y * sc;
    const double t = fma(xs, xs, ys * ys);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    r = fma(r, fma(-t, r, 1.0), r);
    r = fma(r, fma(-t, r, 1.0), r);
    r *= sc;
    return mk(xs * r, -(ys * r));
}
#define ZID_NB 16
#define ZID_NMAX 104
struct zid_slot { cd row[ZID_NB]; cd d; unsigned long long key; int idx; int pad; };
template <int PW> __device__ __forceinline__ void zid_bar_panel() { asm volatile("bar.sync 1, %0;" :: "n"(32 * PW) : "memory"); }
// panel factorisation (warps 0 .. PW-1, thread <-> row): block column k0 .. k0+nbk of As becomes the Gauss-Jordan block column P'
template <int PW, int NB>
__device__ __forceinline__ void zid_panel(cd* As, int lda, int np, int n, int k0, int nbk, zid_slot* slots, int* piv, int& bad, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const bool rowok = tid < np;
    cd p[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) p[j] = (rowok && j < nbk) ? As[tid * lda + k0 + j] : mk(0.0, 0.0);
    // speculative per-row quantities of the NEXT pivot column: 1 / entry and the ordering key of |re| + |im|
    // (bits of a non-negative double order like an unsigned integer; +1 so an eligible zero beats an ineligible row)
    cd dmine = kh_crecip_fast(p[0]);
    unsigned long long key = (rowok && tid >= k0) ? (unsigned long long)__double_as_longlong(cabs1(p[0])) + 1ull : 0ull;
    // The register panel is ROTATED by one column per pivot (the pivot column is always p[0], the next one p[1]) so that the
    // step is a real loop of ~600 instructions instead of 16 unrolled copies: the unrolled form ran out of the instruction
    // cache with a single warp per scheduler and nothing to hide the fetches.
    // pivots k >= n sit in the identity padding (row k = e_k, column k zero elsewhere): their Gauss-Jordan step is the identity,
    // so the loop stops at n and only records them
    const int nsteps = (n - k0 < nbk) ? ((n - k0 > 0) ? n - k0 : 0) : nbk;
    if (tid < nbk - nsteps) piv[k0 + nsteps + tid] = k0 + nsteps + tid;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        const int k = k0 + s;
        zid_slot* sl = slots + (s & 1) * (PW + 1);     // PW warp candidates + the old row k
        // warp-level argmax (two 32-bit reductions + ballot, ties -> smallest row); the warp's winner publishes its row
        const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, khi == mhi ? klo : 0u);
        const unsigned who = __ballot_sync(0xffffffffu, khi == mhi && klo == mlo);
        const bool iswin = lane == __ffs(who) - 1;
        if (iswin || tid == k) {                       // one divergent block for both publishers (row k's copy is only read when pr != k)
            zid_slot* dst = iswin ? sl + warp : sl + PW;
#pragma unroll
            for (int j = 0; j < NB; ++j) dst->row[j] = p[j];
            dst->d = dmine; dst->key = key; dst->idx = tid;
        }
        zid_bar_panel<PW>();
        unsigned long long bk = sl[0].key; int wbest = 0;
#pragma unroll
        for (int w = 1; w < PW; ++w) { const unsigned long long ok = sl[w].key; if (ok > bk) { bk = ok; wbest = w; } }
        const zid_slot* win = sl + wbest;
        const int pr = win->idx;
        const cd d = win->d;
        if (tid == 0) piv[k] = pr;
        if (tid == pr && pr != k) {                    // row k's old values: in its warp's slot if it won there, else in slot 4
            const zid_slot* rk = (sl[k >> 5].idx == k) ? sl + (k >> 5) : sl + PW;
#pragma unroll
            for (int j = 0; j < NB; ++j) p[j] = rk->row[j];
        }
        const cd pv = win->row[0];
        if (pv.x == 0.0 && pv.y == 0.0 && !bad && k < n) bad = k + 1;
        // The row after this step, rotated in place (p[j-1] <- column j).  Row k itself (it becomes the scaled pivot row,
        // row[j] d, and the new last column d) runs through the SAME update as every other row with p = 0 and g = -d:
        // 0 - (-d) row[j] = d row[j], -g = d.  No divergent branch on the warp that owns row k, no copy of the rotated row.
        const bool isk = (tid == k);
        const cd g = isk ? mk(-d.x, -d.y) : p[0] * d;
        cd nx = isk ? mk(0.0, 0.0) : p[1];
        cfms(nx, g, win->row[1]);                      // next pivot column first, its reciprocal overlaps the rest of the update
        dmine = kh_crecip_fast(nx);
        key = (rowok && tid > k) ? (unsigned long long)__double_as_longlong(cabs1(nx)) + 1ull : 0ull;
        p[0] = nx;
#pragma unroll
        for (int j = 2; j < NB; ++j) { cd v = isk ? mk(0.0, 0.0) : p[j]; cfms(v, g, win->row[j]); p[j - 1] = v; }
        p[NB - 1] = mk(-g.x, -g.y);
    }
    if (rowok) {                                       // p[j] holds panel column (j + nsteps) mod NB
#pragma unroll
        for (int j = 0; j < NB; ++j) { const int col = (j + nsteps) & (NB - 1); if (col < nbk) As[tid * lda + k0 + col] = p[j]; }
    }
}
// A[:, tile columns outside [skip0, skip1) or inside [only0, only1)] += P' R on 8x8 DMMA tiles; worker w of nwork takes tiles w, w + nwork, ...
__device__ __forceinline__ void zid_update(cd* As, const cd* R, int lda, int ldr, int np, int k0, int nbk, int c0, int c1, bool inside,
                                           int w, int nwork, int lane) {
    const int lr = lane >> 2, lk = lane & 3;
    const int nstrip = np >> 3, ncols = inside ? (c1 - c0) : nstrip - (c1 - c0), total = nstrip * ncols;
    if (ncols <= 0) return;
    int strip = w / ncols, tcr = w - strip * ncols;               // (one division per call; the tile index then advances incrementally)
    for (int t = w; t < total; t += nwork) {
        int tc = tcr;
        if (inside) tc += c0; else if (tc >= c0) tc += c1 - c0;
        cd* cp = As + (strip * 8 + lr) * lda + tc * 8 + 2 * lk;
        const cd* ap = As + (strip * 8 + lr) * lda + k0 + lk;
        const cd* bp = R + lk * ldr + tc * 8 + lr;
        const cd c0v = cp[0], c1v = cp[1];
        double cr0 = c0v.x, cr1 = c1v.x, ci0 = c0v.y, ci1 = c1v.y;
        if (nbk == ZID_NB) {
            cd av[4], bv[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { av[kk] = ap[kk * 4]; bv[kk] = bp[kk * 4 * ldr]; }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                kh_dmma(cr0, cr1, av[kk].x, bv[kk].x); kh_dmma(ci0, ci1, av[kk].x, bv[kk].y);
                kh_dmma(cr0, cr1, -av[kk].y, bv[kk].y); kh_dmma(ci0, ci1, av[kk].y, bv[kk].x);
            }
        } else {
            for (int kk = 0; kk < (nbk >> 2); ++kk) {
                const cd av = ap[kk * 4], bv = bp[kk * 4 * ldr];
                kh_dmma(cr0, cr1, av.x, bv.x); kh_dmma(ci0, ci1, av.x, bv.y);
                kh_dmma(cr0, cr1, -av.y, bv.y); kh_dmma(ci0, ci1, av.y, bv.x);
            }
        }
        cp[0] = mk(cr0, ci0); cp[1] = mk(cr1, ci1);
        tcr += nwork;
        while (tcr >= ncols) { tcr -= ncols; ++strip; }
    }
}
// GLOBAL = false: the matrix lives in shared memory (n <= 104).  GLOBAL = true: the same single-launch algorithm with the working
// copy in a global-memory scratch (L2 resident: a batch below one wave is a few tens of MB), n <= 256, 8 panel warps -- for SMALL
// batches of the 9x9 ... 11x11 bases (field maps, scalar solves), where the blocked multi-launch variant pays 19 launch latencies
// of ~0.1 ms per inverse with the GPU almost empty.
template <int NW, int PW, bool GLOBAL, int NB = ZID_NB>
__device__ __forceinline__ void zinv_dmma_body_t(const Cta& c, const zinv_args& a) {
    const int n = a.n, b = c.bx, tid = c.tid, warp = tid >> 5, lane = tid & 31;
    const int np = (n + 7) & ~7, lda = GLOBAL ? np : np + 4, ldr = np + 2;
    const cd* A = mat_ptr(a.A, b);
    cd* Out = mat_ptr(a.Ainv, b);
    cd* As = GLOBAL ? a.gwork + (long long)b * np * np : (cd*)KH_SMEM(c);      // [np][lda]
    cd* R = GLOBAL ? (cd*)KH_SMEM(c) : As + np * lda;                          // [NB][ldr]
    zid_slot* slots = (zid_slot*)(R + NB * ldr);   // [2 parities][PW warp candidates + old row k]
    int* piv = (int*)(slots + 2 * (PW + 1));      // [np]
    unsigned long long* bar = (unsigned long long*)(piv + ((np + 1) & ~1));     // mbarrier of the staging copy
    if (GLOBAL) {
        for (int e = tid; e < np * np; e += 32 * NW) { const int i = e / np, j = e - i * np; As[e] = (i < n && j < n) ? A[(long long)i * a.A.ld + j] : mk(i == j ? 1.0 : 0.0, 0.0); }
        __syncthreads();
    } else {
        // identity padding by the threads while the TMA engine stages the n rows (cp.async.bulk, one row each, kh_stage_rows)
        const int padc = np - n, nbot = padc * np;          // bottom rows n..np-1 in full, then columns n..np-1 of the rows above
        for (int e = tid; e < nbot + n * padc; e += 32 * NW) {
            const int i = e < nbot ? n + e / np : (e - nbot) / padc, j = e < nbot ? e % np : n + (e - nbot) % padc;
            As[i * lda + j] = mk(i == j ? 1.0 : 0.0, 0.0);
        }
        kh_stage_rows(c, As, lda, A, a.A.ld, n, n, bar);
    }
    int bad = 0;
    if (warp < PW) zid_panel<PW, NB>(As, lda, np, n, 0, min(NB, np), slots, piv, bad, tid);
    __syncthreads();
    for (int k0 = 0; k0 < np; k0 += NB) {
        const int nbk = min(NB, np - k0);         // NB, or 8 at the end
        const int k1 = k0 + nbk, nbk1 = min(NB, np - k1);      // next panel (nbk1 <= 0: none)
        // ---------------- interchanges of panel k0 on the other columns, pivot rows -> R (zeroed in place): thread <-> column
        {
            const int j = tid - 32 * PW;
            if (j >= 0 && j < np && (j < k0 || j >= k1)) {
                for (int s = 0; s < nbk; ++s) {
                    const int k = k0 + s, pr = piv[k];
                    if (pr != k) { const cd x = As[k * lda + j]; As[k * lda + j] = As[pr * lda + j]; As[pr * lda + j] = x; }
                }
                for (int s = 0; s < nbk; ++s) { R[s * ldr + j] = As[(k0 + s) * lda + j]; As[(k0 + s) * lda + j] = mk(0.0, 0.0); }
            }
        }
        __syncthreads();
        zid_update(As, R, lda, ldr, np, k0, nbk, k0 >> 3, k1 >> 3, false, warp, NW, lane);
        __syncthreads();
        // (look-ahead schedules -- next panel factorised while the other warps finish this update -- were measured three times, with
        //  the panel warps spread over all four schedulers, packed on one, and on two with the update on the rest: no gain,
        //  profiles/r02_experiments.md)
        if (nbk1 > 0 && warp < PW) zid_panel<PW, NB>(As, lda, np, n, k1, nbk1, slots, piv, bad, tid);
        __syncthreads();
    }
    // undo the row interchanges as column interchanges (reverse order): thread j follows stored column j to its final position
    int* dest = (int*)R;
    if (tid < n) {
        int pos = tid;
        for (int k = np - 1; k >= 0; --k) { const int pr = piv[k]; pos = (pos == k) ? pr : ((pos == pr) ? k : pos); }
        dest[tid] = pos;
    }
    __syncthreads();
    for (int i = warp; i < n; i += NW)
        for (int j = lane; j < n; j += 32) Out[(long long)i * a.Ainv.ld + dest[j]] = As[i * lda + j];
    if (tid == 0) zinv_note(a.info, a.info_mode, b, bad);
}
__device__ __forceinline__ void zinv_dmma_body(const Cta& c, const zinv_args& a) { zinv_dmma_body_t<16, 4, false>(c, a); }
// n <= 64 with panels of 8 pivots: half the panel registers and a third less shared memory, so THREE matrices share an SM and
// their latency-bound panels overlap (the 5x5 basis: n = 50)
__device__ __forceinline__ void zinv_dmma8n_body(const Cta& c, const zinv_args& a) { zinv_dmma_body_t<8, 2, false, 8>(c, a); }
static inline size_t zinv_dmma8n_smem(int n) {
    const int np = (n + 7) & ~7;
    return ((size_t)np * (np + 4) + (size_t)8 * (np + 2)) * sizeof(cd) + 6 * sizeof(zid_slot) + (size_t)np * 4 + 32;
}
__device__ __forceinline__ void zinv_l2_body(const Cta& c, const zinv_args& a) { zinv_dmma_body_t<16, 8, true>(c, a); }
#define ZIL_NMAX 256
// ---------------------------------------------------------------------------------------------
// Thread-block-CLUSTER variant for a HANDFUL of matrices beyond one SM's shared memory (104 < n <= 256: the
// dependent inverses of a field-map solve, scalar solves of the 9x9 ... 11x11 bases; batch x CS <= 148, one wave).  A cluster of CS = 4 or 8 CTAs holds ONE
// matrix in its distributed shared memory, CTA r the 16-column blocks [r bpc, (r + 1) bpc): the whole inversion runs out of shared
// memory like zinv_dmma, where the single-CTA zinv_l2 works on an L2-resident copy.  Per block column:
//   owner CTA : register-resident panel (zid_panel) on its own columns;
//   cluster barrier; every CTA PULLS the finished block column P' (np x 16) and its pivot rows from the owner's shared memory
//   (ld.shared::cluster through cooperative_groups' map_shared_rank) into a local buffer;
//   every CTA : interchanges, pivot rows -> R, DMMA rank-16 update of its own columns (P' and C from local shared memory).
// One cluster barrier per block column: the owner's panel columns are not touched again before the next one.
#include <cooperative_groups.h>
__device__ __forceinline__ void zidc_update(cd* Al, const cd* Pb, const cd* R, int ldc, int ldp, int ldr, int np, int nbk, int ntc, int s0, int s1,
                                            int w, int nwork, int lane) {
    // local tile columns [0, ntc) except [s0, s1) (the panel's own columns on the owner; s0 = s1 elsewhere)
    const int lr = lane >> 2, lk = lane & 3;
    const int nstrip = np >> 3, ncols = ntc - (s1 - s0), total = nstrip * ncols;
    if (ncols <= 0) return;
    int strip = w / ncols, tcr = w - strip * ncols;
    for (int t = w; t < total; t += nwork) {
        const int tc = tcr >= s0 ? tcr + (s1 - s0) : tcr;
        cd* cp = Al + (strip * 8 + lr) * ldc + tc * 8 + 2 * lk;
        const cd* ap = Pb + (strip * 8 + lr) * ldp + lk;
        const cd* bp = R + lk * ldr + tc * 8 + lr;
        const cd c0v = cp[0], c1v = cp[1];
        double cr0 = c0v.x, cr1 = c1v.x, ci0 = c0v.y, ci1 = c1v.y;
        for (int kk = 0; kk < (nbk >> 2); ++kk) {
            const cd av = ap[kk * 4], bv = bp[kk * 4 * ldr];
            kh_dmma(cr0, cr1, av.x, bv.x); kh_dmma(ci0, ci1, av.x, bv.y);
            kh_dmma(cr0, cr1, -av.y, bv.y); kh_dmma(ci0, ci1, av.y, bv.x);
        }
        cp[0] = mk(cr0, ci0); cp[1] = mk(cr1, ci1);
        tcr += nwork;
        while (tcr >= ncols) { tcr -= ncols; ++strip; }
    }
}
struct zidc_shape { int np, nblk, bpc, wc, ldc, ldp, ldr; size_t smem; };
static inline __host__ __device__ zidc_shape zidc_make(int n, int CS) {
    zidc_shape h;
    h.np = (n + 7) & ~7; h.nblk = (h.np + ZID_NB - 1) / ZID_NB; h.bpc = (h.nblk + CS - 1) / CS; h.wc = h.bpc * ZID_NB;
    h.ldc = h.wc + 4; h.ldr = h.wc + 2;
    h.ldp = ZID_NB + 4;
    auto bytes = [&](int ldp) { return (size_t)h.np * h.ldc * sizeof(cd) + (size_t)h.np * ldp * sizeof(cd) + (size_t)ZID_NB * h.ldr * sizeof(cd) +
                                       18 * sizeof(zid_slot) + (size_t)(h.np + 2) * 4 + 64; };
    if (bytes(h.ldp) > (size_t)227 * 1024) h.ldp = ZID_NB + 2;       // (2-way conflicts on the P' fragments instead of none)
    h.smem = bytes(h.ldp);
    return h;
}
template <int CS>
__device__ __forceinline__ void zinv_cluster_body_t(const Cta& c, const zinv_args& a) {
    namespace cg = cooperative_groups;
    constexpr int NW = 16, PW = 8, NB = ZID_NB;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int n = a.n, b = c.bx / CS, tid = c.tid, warp = tid >> 5, lane = tid & 31;
    const zidc_shape h = zidc_make(n, CS);
    const int np = h.np, ldc = h.ldc, ldp = h.ldp, ldr = h.ldr, wc = h.wc;
    const int c0 = rank * wc, c1 = min(np, c0 + wc), ncl = max(0, c1 - c0);        // this CTA's global columns [c0, c1)
    const cd* A = mat_ptr(a.A, b);
    cd* Out = mat_ptr(a.Ainv, b);
    cd* Al = (cd*)KH_SMEM(c);                       // [np][ldc]   own columns
    cd* Pb = Al + np * ldc;                         // [np][ldp]   block column P' of the current step
    cd* R = Pb + np * ldp;                          // [NB][ldr]   pivot rows, own columns
    zid_slot* slots = (zid_slot*)(R + NB * ldr);
    int* piv = (int*)(slots + 2 * (PW + 1));        // [np] (every CTA collects all pivots)
    unsigned long long* bar = (unsigned long long*)(piv + ((np + 2) & ~1));
    if (rank == 0 && tid == 0 && a.info && a.info_mode == 0) a.info[b] = 0;
    {   // own columns: rows of the matrix by TMA bulk copies, identity padding by the threads
        const int ncv = min(c1, n) - c0;                                       // columns that exist in the matrix
        for (int e = tid; e < np * ncl; e += 32 * NW) {
            const int i = e / ncl, j = e - i * ncl;
            if (i >= n || c0 + j >= n) Al[i * ldc + j] = mk(i == c0 + j ? 1.0 : 0.0, 0.0);
        }
        if (ncv > 0) kh_stage_rows(c, Al, ldc, A + c0, a.A.ld, n, ncv, bar);
        else __syncthreads();
    }
    int bad = 0;
    for (int kb = 0; kb < h.nblk; ++kb) {
        const int k0 = kb * NB, nbk = min(NB, np - k0), owner = kb / h.bpc;
        if (rank == owner) {
            if (warp < PW) zid_panel<PW, NB>(Al - c0, ldc, np, n, k0, nbk, slots, piv, bad, tid);
        }
        cl.sync();                                  // the owner's block column is final; every CTA is past the previous update
        {
            const cd* src = cl.map_shared_rank(Al, owner) + (k0 - owner * wc);
            const int* psrc = cl.map_shared_rank(piv, owner);
            for (int e = tid; e < np * nbk; e += 32 * NW) { const int i = e / nbk, j = e - i * nbk; Pb[i * ldp + j] = src[i * ldc + j]; }
            if (rank != owner && tid < nbk) piv[k0 + tid] = psrc[k0 + tid];
        }
        __syncthreads();
        const int p0 = rank == owner ? k0 - c0 : 0, p1 = rank == owner ? p0 + nbk : 0;      // local columns of the panel (owner only)
        if (tid < ncl && (tid < p0 || tid >= p1)) {     // interchanges on the own columns, pivot rows -> R (zeroed in place): thread <-> column
            const int j = tid;
            for (int s = 0; s < nbk; ++s) {
                const int k = k0 + s, pr = piv[k];
                if (pr != k) { const cd x = Al[k * ldc + j]; Al[k * ldc + j] = Al[pr * ldc + j]; Al[pr * ldc + j] = x; }
            }
            for (int s = 0; s < nbk; ++s) { R[s * ldr + j] = Al[(k0 + s) * ldc + j]; Al[(k0 + s) * ldc + j] = mk(0.0, 0.0); }
        }
        __syncthreads();
        zidc_update(Al, Pb, R, ldc, ldp, ldr, np, nbk, ncl >> 3, p0 >> 3, p1 >> 3, warp, NW, lane);
        __syncthreads();
    }
    cl.sync();                                      // (nobody leaves while a peer may still read its shared memory)
    // undo the row interchanges as column interchanges: stored global column c0 + j goes to column dest
    int* dest = (int*)R;
    if (tid < ncl) {
        int pos = c0 + tid;
        for (int k = np - 1; k >= 0; --k) { const int pr = piv[k]; pos = (pos == k) ? pr : ((pos == pr) ? k : pos); }
        dest[tid] = pos;
    }
    __syncthreads();
    const int ncv = min(c1, n) - c0;
    for (int i = warp; i < n; i += NW)
        for (int j = lane; j < ncv; j += 32) Out[(long long)i * a.Ainv.ld + dest[j]] = Al[i * ldc + j];
    if (tid == 0 && bad && a.info) {
        if (a.info_mode > 0) KH_ATOMIC_OR(&a.info[b / a.info_mode], 2);
        else KH_ATOMIC_MAX(&a.info[b], bad);
    }
}
__device__ __forceinline__ void zinv_cluster4_body(const Cta& c, const zinv_args& a) { zinv_cluster_body_t<4>(c, a); }
__device__ __forceinline__ void zinv_cluster8_body(const Cta& c, const zinv_args& a) { zinv_cluster_body_t<8>(c, a); }
static inline long long zinv_l2_work_cd(int n) { const long long np = (n + 7) & ~7; return np * np; }
static inline size_t zinv_l2_smem(int n) {
    const int np = (n + 7) & ~7;
    return (size_t)ZID_NB * (np + 2) * sizeof(cd) + 18 * sizeof(zid_slot) + (size_t)np * 4 + 32;
}
static inline size_t zinv_dmma_smem(int n) {
    const int np = (n + 7) & ~7;
    return ((size_t)np * (np + 4) + (size_t)ZID_NB * (np + 2)) * sizeof(cd) + 10 * sizeof(zid_slot) + (size_t)np * 4 + 32;
}
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
    int info_mode = 0;
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
    if (a.info && c.tid == 0) {
        if (a.info_mode > 0) { if (bad) KH_ATOMIC_OR(&a.info[b / a.info_mode], 2); }
        else if (k0 == 0) a.info[b] = bad; else if (bad && a.info[b] == 0) a.info[b] = bad;
    }
}

#ifndef KH_HOST_EMU
// Block column of 32 pivots of the blocked variant for n <= 512, built on the register-resident panel of the shared-memory kernel
// (zid_panel: thread <-> row, 16 panel entries per row in registers, one named barrier per pivot) working directly on the L2-resident
// matrix: ~1.2 k cycles per pivot where the shared-memory panel above (five CTA barriers and an n x nb sweep per pivot) took ~15 k at
// n = 242.  Two half panels A, B of 16 columns; the rank-16 update of A is applied inside the kernel only where B needs it (B's
// columns, and B's pivot rows R_B), so that the batched GEMM afterwards applies BOTH halves to the rest of the matrix in one pass
// over it (K = 32: the update is HBM bound, n^2 32 B per pass):
//   1 panel A   2 interchanges of A on all other columns, pivot rows -> R_A (zeroed in place)   3 B's columns += P'_A R_A
//   4 panel B   5 interchanges of B (A's columns included)   6 rows of B's pivots -> R_B (zeroed in place, A's columns too)
//   7 A's columns += P'_B R_B[:, A's columns]: the 32 columns are now the transformation of the whole block
template <int PW>
__device__ __forceinline__ void zinvb_panel32_body_t(const Cta& c, const zinvb_args& a) {
    const int n = a.n, b = c.bx, k0 = a.k0, tid = c.tid, warp = tid >> 5;
    const int nb = (32 < n - k0) ? 32 : n - k0, nA = nb < ZID_NB ? nb : ZID_NB, nB = nb - nA, kB = k0 + ZID_NB;
    cd* A = mat_ptr(a.A, b);
    const int ld = a.A.ld;
    cd* R = a.R + (long long)b * a.r_stride;          // [32][n]
    int* pivg = a.piv + (long long)b * a.piv_stride;
    zid_slot* slots = (zid_slot*)KH_SMEM(c);          // [2 parities][PW warp candidates + old row k]
    cd* T16 = (cd*)(slots + 2 * (PW + 1));            // [16][16] staging of a small operand
    int* pivs = (int*)(T16 + ZID_NB * ZID_NB);        // [32]
    int* piv = pivs - k0;
    int bad = 0;
    auto interchange = [&](int p0, int np_, int c0, int c1) {     // row swaps of the pivots p0 .. p0+np_ on the columns outside [c0, c1)
        for (int j = tid; j < n; j += c.nthr) {
            if (j >= c0 && j < c1) continue;
            for (int s = 0; s < np_; ++s) {
                const int k = p0 + s, pr = piv[k];
                if (pr != k) { const cd x = A[(long long)k * ld + j]; A[(long long)k * ld + j] = A[(long long)pr * ld + j]; A[(long long)pr * ld + j] = x; }
            }
        }
    };
    // ---- 1, 2
    if (warp < PW) zid_panel<PW, ZID_NB>(A, ld, n, n, k0, nA, slots, piv, bad, tid);
    __syncthreads();
    interchange(k0, nA, k0, k0 + nA);
    __syncthreads();
    for (int j = tid; j < n; j += c.nthr) {
        if (j >= k0 && j < k0 + nA) continue;
        for (int s = 0; s < nA; ++s) { R[(long long)s * n + j] = A[(long long)(k0 + s) * ld + j]; A[(long long)(k0 + s) * ld + j] = mk(0.0, 0.0); }
    }
    __syncthreads();
    if (nB > 0) {
        // ---- 3: B's columns += P'_A R_A[:, B]
        for (int e = tid; e < nA * nB; e += c.nthr) { const int t = e / nB, jb = e - t * nB; T16[t * ZID_NB + jb] = R[(long long)t * n + kB + jb]; }
        __syncthreads();
        for (int i = tid; i < n; i += c.nthr) {
            cd acc[ZID_NB];
#pragma unroll
            for (int jb = 0; jb < ZID_NB; ++jb) acc[jb] = mk(0.0, 0.0);
            for (int t = 0; t < nA; ++t) {
                const cd p = A[(long long)i * ld + k0 + t];
#pragma unroll
                for (int jb = 0; jb < ZID_NB; ++jb) cfma(acc[jb], p, T16[t * ZID_NB + jb]);
            }
#pragma unroll
            for (int jb = 0; jb < ZID_NB; ++jb) if (jb < nB) { cd* q = &A[(long long)i * ld + kB + jb]; *q = *q + acc[jb]; }
        }
        __syncthreads();
        // ---- 4, 5
        if (warp < PW) zid_panel<PW, ZID_NB>(A, ld, n, n, kB, nB, slots, piv, bad, tid);
        __syncthreads();
        interchange(kB, nB, kB, kB + nB);
        __syncthreads();
        // ---- 6: R_B = the rows of B's pivots as they stand (zeroed in place).  NOT updated by A's pending rank-16 term: after step 7
        // the 32 columns hold the Gauss-Jordan transformation of the WHOLE block, [F_A F_B] with F_A = P'_A(rows of B zeroed) +
        // P'_B P'_A[rows of B], and  M + F_A R_A + F_B R_B(raw)  is exactly the two sequential rank-16 updates.
        for (int j = tid; j < n; j += c.nthr) {
            if (j >= kB && j < kB + nB) continue;
            for (int s_ = 0; s_ < nB; ++s_) {
                R[(long long)(ZID_NB + s_) * n + j] = A[(long long)(kB + s_) * ld + j];
                A[(long long)(kB + s_) * ld + j] = mk(0.0, 0.0);
            }
        }
        __syncthreads();
        // ---- 7: A's columns += P'_B R_B[:, A]
        for (int e = tid; e < nB * nA; e += c.nthr) { const int s_ = e / nA, ja = e - s_ * nA; T16[s_ * ZID_NB + ja] = R[(long long)(ZID_NB + s_) * n + k0 + ja]; }
        __syncthreads();
        for (int i = tid; i < n; i += c.nthr) {
            cd acc[ZID_NB];
#pragma unroll
            for (int ja = 0; ja < ZID_NB; ++ja) acc[ja] = mk(0.0, 0.0);
            for (int s_ = 0; s_ < nB; ++s_) {
                const cd p = A[(long long)i * ld + kB + s_];
#pragma unroll
                for (int ja = 0; ja < ZID_NB; ++ja) cfma(acc[ja], p, T16[s_ * ZID_NB + ja]);
            }
#pragma unroll
            for (int ja = 0; ja < ZID_NB; ++ja) if (ja < nA) { cd* q = &A[(long long)i * ld + k0 + ja]; *q = *q + acc[ja]; }
        }
    }
    __syncthreads();
    if (tid < nb) pivg[k0 + tid] = pivs[tid];
    if (a.info && tid == 0) {
        if (a.info_mode > 0) { if (bad) KH_ATOMIC_OR(&a.info[b / a.info_mode], 2); }
        else if (k0 == 0) a.info[b] = bad; else if (bad && a.info[b] == 0) a.info[b] = bad;
    }
}
__device__ __forceinline__ void zinvb_panel32a_body(const Cta& c, const zinvb_args& a) { zinvb_panel32_body_t<8>(c, a); }      // n <= 256
__device__ __forceinline__ void zinvb_panel32b_body(const Cta& c, const zinvb_args& a) { zinvb_panel32_body_t<16>(c, a); }     // n <= 512
#endif

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
static inline long long zinv_work_cd(int n) { const int nb = zinvb_nb(n) > 32 ? zinvb_nb(n) : 32; return (long long)nb * n + (n + 3) / 4 + 4; }
#ifndef KH_ZINV_BLOCKED_MIN
#define KH_ZINV_BLOCKED_MIN 101
#endif

static inline int zinv_blocked_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info, cd* work, int info_mode = 0) {
    int e;
    if (A.p != Ainv.p) {
        zcopym_args cp{n, A, Ainv};
        if ((e = kh_launch<zcopym_args, zcopym_body>(dim3(batch), 256, 0, st, cp, "zinv", 0.0))) return e;
    }
#ifndef KH_HOST_EMU
    const bool fastpanel = n <= 512;                   // register-resident block column of 32 pivots on the L2-resident matrix
#else
    const bool fastpanel = false;
#endif
    const int nb = fastpanel ? 32 : zinvb_nb(n);
    const long long wstride = zinv_work_cd(n);
    zinvb_args a;
    a.n = n; a.nb = nb; a.A = Ainv; a.R = work; a.r_stride = wstride; a.info = info; a.lds = nb + 1; a.info_mode = info_mode;
    a.piv = (int*)(work + (long long)nb * n); a.piv_stride = wstride * 4;
    const size_t sm = (size_t)n * 16 + 32 * 16 + 128 * 8 + 32 * 4 + 16 + (size_t)n * a.lds * sizeof(cd);
    for (int k0 = 0; k0 < n; k0 += nb) {
        a.k0 = k0;
        const int nbk = nb < n - k0 ? nb : n - k0;
#ifndef KH_HOST_EMU
        if (fastpanel) {
            const size_t psm = (size_t)34 * sizeof(zid_slot) + (size_t)ZID_NB * ZID_NB * sizeof(cd) + 64 * sizeof(int) + 16;
            if (n <= 256) e = kh_launch<zinvb_args, zinvb_panel32a_body, 512, 1>(dim3(batch), 512, psm, st, a, "zinv", 0.0);
            else e = kh_launch<zinvb_args, zinvb_panel32b_body, 512, 1>(dim3(batch), 512, psm, st, a, "zinv", 0.0);
            if (e) return e;
        } else
#endif
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

static inline int zinv_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef Ainv, int* info, cd* work = nullptr, long long work_cd = 0, int info_mode = 0) {
    if (batch <= 0 || n <= 0) return 0;
#ifndef KH_HOST_EMU
    {   // a handful of mid-size matrices: one thread-block cluster per matrix, the matrix in distributed shared memory
        const char* e = getenv("KH_ZINV_CLUSTER_MAXCTAS");   // (test / tuning switch; 0 disables)
        const int maxctas = e ? atoi(e) : 148;             // one wave of clusters; beyond it zinv_l2 (one CTA per matrix) is as fast
        if (n > ZID_NMAX && n <= ZIL_NMAX) {
            const int CS = zidc_make(n, 4).smem <= (size_t)KH_SMEM_MAX ? 4 : 8;
            const zidc_shape h = zidc_make(n, CS);
            if (h.smem <= (size_t)KH_SMEM_MAX && (long long)batch * CS <= maxctas) {
                zinv_args g;
                g.n = n; g.A = A; g.Ainv = Ainv; g.info = info; g.info_mode = info_mode; g.use_smem = 0; g.ld_s = 0;
                if (CS == 4) return kh_launch_cluster<zinv_args, zinv_cluster4_body, 512, 1>(dim3(batch * 4), 512, h.smem, 4, st, g, "zinv", 8.0 * n * n * n * batch);
                return kh_launch_cluster<zinv_args, zinv_cluster8_body, 512, 1>(dim3(batch * 8), 512, h.smem, 8, st, g, "zinv", 8.0 * n * n * n * batch);
            }
        }
    }
    {   // small batches of mid-size matrices: ONE launch, working copy in L2 (see zinv_dmma_body_t<.., true>)
        const char* e = getenv("KH_ZINV_L2_MAXBATCH");       // (test / tuning switch; default: two waves of one CTA per SM)
        const int l2max = e ? atoi(e) : 296;
        if (n > ZID_NMAX && n <= ZIL_NMAX && batch <= l2max && work && work_cd >= (long long)batch * zinv_l2_work_cd(n)) {
            zinv_args g;
            g.n = n; g.A = A; g.Ainv = Ainv; g.info = info; g.info_mode = info_mode; g.use_smem = 0; g.ld_s = 0; g.gwork = work;
            return kh_launch<zinv_args, zinv_l2_body, 512, 1>(dim3(batch), 512, zinv_l2_smem(n), st, g, "zinv", 8.0 * n * n * n * batch);
        }
    }
#endif
    if (n >= KH_ZINV_BLOCKED_MIN && work && work_cd >= (long long)batch * zinv_work_cd(n)) return zinv_blocked_launch(st, batch, n, A, Ainv, info, work, info_mode);
    zinv_args a;
    a.n = n; a.A = A; a.Ainv = Ainv; a.info = info; a.info_mode = info_mode;
#ifndef KH_HOST_EMU
    if (n <= 64 && n >= 16)          // panels of 8, three CTAs per SM: 5.1 -> 7.6 TFLOP/s at n = 50 against panels of 16 with two
        return kh_launch<zinv_args, zinv_dmma8n_body, 256, 3>(dim3(batch), 256, zinv_dmma8n_smem(n), st, a, "zinv", 8.0 * n * n * n * batch);
    if (n <= ZID_NMAX && n >= 16)
        return kh_launch<zinv_args, zinv_dmma_body, 512, 1>(dim3(batch), 512, zinv_dmma_smem(n), st, a, "zinv", 8.0 * n * n * n * batch);
#endif
    a.ld_s = n | 1;
    a.use_smem = zinv_smem_bytes(n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int threads = n <= 64 ? 256 : 512;
    return kh_launch<zinv_args, zinv_body>(dim3(batch), threads, zinv_smem_bytes(n, a.ld_s, a.use_smem), st, a, "zinv", 8.0 * n * n * n * batch);
}
