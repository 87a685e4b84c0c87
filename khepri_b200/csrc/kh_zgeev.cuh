// Batched non-Hermitian complex128 eigensolver, one CTA per matrix (replaces numpy.linalg.eig /
// LAPACK zgeev at khepri/alternative.py:172 for the Omega^2 = P Q problem of every patterned layer).
//
// Pipeline (all inside one kernel, matrix resident in shared memory when n <= ~118):
//   1. diagonal balancing with powers of two (exact similarity, as zgebal 'S')
//   2. Householder reduction to upper Hessenberg form, Schur vectors accumulated (Zt = Z^T in HBM,
//      so that every column operation on Z is a coalesced row operation on Zt)
//   3. single-shift QR iteration with Wilkinson / exceptional shifts and LAPACK's (zlahqr)
//      deflation test, Givens rotations applied to full rows/columns (Schur form T is needed)
//   4. eigenvectors of T by back substitution (as ztrevc), one thread per eigenvector
// The caller finishes with one DMMA GEMM:  W = diag(scale) * Z * X.
// Eigenvalue order / eigenvector scaling are free: parity is on physical outputs.
#pragma once
#include "kh_common.cuh"

struct zgeev_args {
    int n;
    MatRef A;          // input (not modified unless it aliases Hw)
    MatRef Hw;         // n x n work matrix in global memory (used when the matrix does not fit in smem)
    MatRef Zt;         // out: transposed Schur vectors
    MatRef X;          // out: eigenvectors of the triangular factor (upper triangular, unit diagonal before normalisation)
    cd* w; long long w_stride;          // out: eigenvalues
    cd* scale; long long scale_stride;  // out: balancing factors as complex (imag 0)
    int* info;         // out: 0 ok, >0 = QR iteration failed to converge at that index+1
    int use_smem, ld_s;
};

#define ZGEEV_EPS 2.220446049250313e-16   /* LAPACK ulp = eps*base */
#define ZGEEV_SAFMIN 2.2250738585072014e-308

struct kh_givens { double c; cd s, r; };
KH_DEV kh_givens make_givens(cd f, cd g) {
    kh_givens G;
    double g2 = cabs2(g);
    if (g2 == 0.0) { G.c = 1.0; G.s = mk(0, 0); G.r = f; return G; }
    double f2 = cabs2(f);
    if (f2 == 0.0) {
        double ga = sqrt(g2);
        G.c = 0.0; G.s = (1.0 / ga) * cconj(g); G.r = mk(ga, 0);
        return G;
    }
    double f1 = sqrt(f2), nrm = sqrt(f2 + g2);
    cd fu = (1.0 / f1) * f;
    G.c = f1 / nrm;
    G.s = (1.0 / nrm) * (fu * cconj(g));
    G.r = nrm * fu;
    return G;
}

KH_DEV void zgeev_body(const Cta& c, const zgeev_args& a) {
    const int n = a.n, b = c.bx;
    const cd* A = mat_ptr(a.A, b);
    cd* Zt = mat_ptr(a.Zt, b);
    cd* X = mat_ptr(a.X, b);
    const int ldz = a.Zt.ld, ldx = a.X.ld;
    cd* wout = a.w + (long long)b * a.w_stride;
    cd* scout = a.scale + (long long)b * a.scale_stride;

    // shared: [vv n][uu n][scratch 128 dbl (+64 int)][dsc n dbl][ctl 8 int][H]
    cd* vv = (cd*)c.smem;
    cd* uu = vv + n;
    double* scratch = (double*)(uu + n);
    double* dsc = scratch + 128;
    int* ctl = (int*)(dsc + n);
    cd* H; int ld;
    if (a.use_smem) { H = (cd*)(((uintptr_t)(ctl + 8) + 15) & ~(uintptr_t)15); ld = a.ld_s; }
    else { H = mat_ptr(a.Hw, b); ld = a.Hw.ld; }
#define HH(i, j) H[(long long)(i) * ld + (j)]
#define ZT(i, j) Zt[(long long)(i) * ldz + (j)]
#define XX(i, j) X[(long long)(i) * ldx + (j)]

    if (a.use_smem || H != A)
        for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; HH(i, j) = A[(long long)i * a.A.ld + j]; }
    for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; ZT(i, j) = mk(i == j ? 1.0 : 0.0, 0.0); }
    for (int i = c.tid; i < n; i += c.nthr) dsc[i] = 1.0;
    c.sync();

    // ---------------------------------------------------------------- 1. balancing
    for (int sweep = 0; sweep < 12; ++sweep) {
        double changed = 0.0;
        for (int i = c.tid; i < n; i += c.nthr) {
            double cn = 0.0, rn = 0.0;
            for (int j = 0; j < n; ++j) if (j != i) { cn += cabs1(HH(j, i)); rn += cabs1(HH(i, j)); }
            double f = 1.0;
            if (cn != 0.0 && rn != 0.0 && cn <= 1e300 && rn <= 1e300) {      // (NaN/inf rows are left alone)
                double g = rn * 0.5, s = cn + rn, cc = cn;
                for (int q = 0; q < 1100 && cc < g; ++q) { f *= 2.0; cc *= 4.0; }
                g = rn * 2.0;
                for (int q = 0; q < 1100 && cc >= g; ++q) { f *= 0.5; cc *= 0.25; }
                if ((cc + rn) / f >= 0.95 * s) f = 1.0;
            }
            uu[i].x = f;
            if (f != 1.0) changed = 1.0;
        }
        changed = cta_max(c, changed, scratch);
        c.sync();
        if (changed == 0.0) break;
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            double f = uu[j].x / uu[i].x;
            if (f != 1.0) HH(i, j) = f * HH(i, j);
        }
        for (int i = c.tid; i < n; i += c.nthr) dsc[i] *= uu[i].x;
        c.sync();
    }
    for (int i = c.tid; i < n; i += c.nthr) scout[i] = mk(dsc[i], 0.0);

    // ---------------------------------------------------------------- 2. Hessenberg reduction
    for (int k = 0; k + 2 < n; ++k) {
        double part = 0.0;
        for (int i = k + 2 + c.tid; i < n; i += c.nthr) part += cabs2(HH(i, k));
        double xn2 = cta_sum(c, part, scratch);
        cd alpha = HH(k + 1, k);
        c.sync();
        if (xn2 == 0.0 && alpha.y == 0.0) continue;          // already reduced: H_k = I
        double beta = -copysign(sqrt(cabs2(alpha) + xn2), alpha.x);
        cd tau = mk((beta - alpha.x) / beta, -alpha.y / beta);
        cd sc = crecip(alpha - mk(beta, 0.0));
        for (int i = k + 1 + c.tid; i < n; i += c.nthr) {
            vv[i] = (i == k + 1) ? mk(1.0, 0.0) : HH(i, k) * sc;
            HH(i, k) = (i == k + 1) ? mk(beta, 0.0) : mk(0.0, 0.0);
        }
        c.sync();
        // left: H[k+1:, k+1:] <- (I - conj(tau) v v^H) H[k+1:, k+1:]
        for (int j = k + 1 + c.tid; j < n; j += c.nthr) {
            cd w0 = mk(0, 0), w1 = mk(0, 0);
            int i = k + 1;
            for (; i + 1 < n; i += 2) { cfma(w0, cconj(vv[i]), HH(i, j)); cfma(w1, cconj(vv[i + 1]), HH(i + 1, j)); }
            if (i < n) cfma(w0, cconj(vv[i]), HH(i, j));
            cd wj = cconj(tau) * (w0 + w1);
            for (i = k + 1; i < n; ++i) cfms(HH(i, j), vv[i], wj);
        }
        c.sync();
        // right: H[:, k+1:] <- H[:, k+1:] (I - tau v v^H)   and the same for Z (rows of Zt)
        for (int r = c.tid; r < 2 * n; r += c.nthr) {
            if (r < n) {
                cd u0 = mk(0, 0), u1 = mk(0, 0);
                int j = k + 1;
                for (; j + 1 < n; j += 2) { cfma(u0, HH(r, j), vv[j]); cfma(u1, HH(r, j + 1), vv[j + 1]); }
                if (j < n) cfma(u0, HH(r, j), vv[j]);
                cd ur = tau * (u0 + u1);
                for (j = k + 1; j < n; ++j) cfms(HH(r, j), ur, cconj(vv[j]));
            } else {
                int i = r - n;
                cd u0 = mk(0, 0), u1 = mk(0, 0);
                int j = k + 1;
                for (; j + 1 < n; j += 2) { cfma(u0, ZT(j, i), vv[j]); cfma(u1, ZT(j + 1, i), vv[j + 1]); }
                if (j < n) cfma(u0, ZT(j, i), vv[j]);
                cd ur = tau * (u0 + u1);
                for (j = k + 1; j < n; ++j) cfms(ZT(j, i), ur, cconj(vv[j]));
            }
        }
        c.sync();
    }

    // ---------------------------------------------------------------- 3. shifted QR iteration
    const double smlnum = ZGEEV_SAFMIN * ((double)n / ZGEEV_EPS);
    const int itmax = 30 * (n > 10 ? n : 10);
    int fail = 0;
    int iact = n - 1, its = 0;
    while (iact >= 0) {
        // locate the active block [l, iact]
        if (c.tid == 0) ctl[0] = 0;
        c.sync();
        for (int k = iact - c.tid; k >= 1; k -= c.nthr) {
            cd hs = HH(k, k - 1);
            bool negl = false;
            if (cabs1(hs) <= smlnum) negl = true;
            else {
                double tst = cabs1(HH(k - 1, k - 1)) + cabs1(HH(k, k));
                if (tst == 0.0) {
                    if (k - 2 >= 0) tst += cabs1(HH(k - 1, k - 2));
                    if (k + 1 <= n - 1) tst += cabs1(HH(k + 1, k));
                }
                if (cabs1(hs) <= ZGEEV_EPS * tst) {
                    double h12 = cabs1(HH(k - 1, k)), h21 = cabs1(hs);
                    double ab = fmax(h21, h12), ba = fmin(h21, h12);
                    double d1 = cabs1(HH(k, k)), d2 = cabs1(HH(k - 1, k - 1) - HH(k, k));
                    double aa = fmax(d1, d2), bb = fmin(d1, d2);
                    double s = aa + ab;
                    if (ba * (ab / s) <= fmax(smlnum, ZGEEV_EPS * (bb * (aa / s)))) negl = true;
                }
            }
            if (negl) { KH_ATOMIC_MAX(&ctl[0], k); break; }
        }
        c.sync();
        const int l = ctl[0];
        c.sync();
        if (l > 0 && c.tid == 0) HH(l, l - 1) = mk(0.0, 0.0);
        if (l >= iact) {                       // one eigenvalue converged
            if (c.tid == 0) wout[iact] = HH(iact, iact);
            iact -= 1; its = 0;
            c.sync();
            continue;
        }
        its += 1;
        if (its > itmax) { fail = iact + 1; break; }
        // shift
        cd t;
        if (its % 10 == 0 && (its / 10) % 2 == 1) t = HH(l, l) + mk(0.75 * cabs1(HH(l + 1, l)), 0.0);
        else if (its % 10 == 0) t = HH(iact, iact) + mk(0.75 * cabs1(HH(iact, iact - 1)), 0.0);
        else {
            t = HH(iact, iact);
            cd u = csqrt_(HH(iact - 1, iact)) * csqrt_(HH(iact, iact - 1));
            double s = cabs1(u);
            if (s != 0.0) {
                cd x = 0.5 * (HH(iact - 1, iact - 1) - t);
                double sx = cabs1(x);
                s = fmax(s, sx);
                cd xs = (1.0 / s) * x, us = (1.0 / s) * u;
                cd y = s * csqrt_(xs * xs + us * us);
                if (sx > 0.0) {
                    cd xn = (1.0 / sx) * x;
                    if (xn.x * y.x + xn.y * y.y < 0.0) y = -y;
                }
                t = t - u * (u / (x + y));
            }
        }
        cd f = HH(l, l) - t, g = HH(l + 1, l);
        c.sync();   // all threads have read H before the sweep starts writing
        // one implicit single-shift QR sweep over the active block (two barriers per rotation)
        for (int k = l; k < iact; ++k) {
            kh_givens G = make_givens(f, g);
            // rows k, k+1 over columns k..n-1
            for (int j = k + c.tid; j < n; j += c.nthr) {
                cd h0 = HH(k, j), h1 = HH(k + 1, j);
                HH(k, j) = G.c * h0 + G.s * h1;
                HH(k + 1, j) = G.c * h1 - cconj(G.s) * h0;
            }
            c.sync();
            // columns k, k+1 over rows 0..min(k+2, iact); the same columns of Z; bulge bookkeeping
            if (k > l && c.tid == 0) { HH(k, k - 1) = G.r; HH(k + 1, k - 1) = mk(0.0, 0.0); }
            const int rmax = (k + 2 < iact) ? k + 2 : iact;
            for (int r = c.tid; r < 2 * n; r += c.nthr) {
                if (r < n) {
                    if (r <= rmax) {
                        cd h0 = HH(r, k), h1 = HH(r, k + 1);
                        HH(r, k) = G.c * h0 + cconj(G.s) * h1;
                        HH(r, k + 1) = G.c * h1 - G.s * h0;
                    }
                } else {
                    int i = r - n;
                    cd z0 = ZT(k, i), z1 = ZT(k + 1, i);
                    ZT(k, i) = G.c * z0 + cconj(G.s) * z1;
                    ZT(k + 1, i) = G.c * z1 - G.s * z0;
                }
            }
            c.sync();
            if (k + 1 < iact) { f = HH(k + 1, k); g = HH(k + 2, k); }   // the bulge for the next rotation
        }
    }
    c.sync();

    // ---------------------------------------------------------------- 4. eigenvectors of T
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int i = e / n, k = e - i * n;
        XX(i, k) = (i < k) ? -HH(i, k) : mk(i == k ? 1.0 : 0.0, 0.0);
    }
    c.sync();
    for (int k = c.tid; k < n; k += c.nthr) {
        cd tkk = HH(k, k);
        double smin = fmax(ZGEEV_EPS * cabs1(tkk), smlnum);
        for (int j = n - 2; j >= 0; --j) {
            if (j < k) {
                cd d = HH(j, j) - tkk;
                if (cabs1(d) < smin) d = mk(smin, 0.0);
                cd xj = XX(j, k) / d;
                XX(j, k) = xj;
                for (int i = 0; i < j; ++i) { cd v = XX(i, k); cfms(v, xj, HH(i, j)); XX(i, k) = v; }
            }
        }
        double emax = 0.0;
        for (int i = 0; i <= k; ++i) emax = fmax(emax, cabs1(XX(i, k)));
        double r = 1.0 / emax;
        for (int i = 0; i <= k; ++i) XX(i, k) = r * XX(i, k);
    }
    if (a.info && c.tid == 0) a.info[b] = fail;
#undef HH
#undef ZT
#undef XX
}

static inline size_t zgeev_smem_bytes(int n, int ld_s, int use_smem) {
    size_t s = (size_t)2 * n * sizeof(cd) + 128 * sizeof(double) + (size_t)n * sizeof(double) + 8 * sizeof(int) + 16;
    if (use_smem) s += (size_t)n * ld_s * sizeof(cd);
    return s;
}

static inline int zgeev_launch(kh_stream_t st, int batch, zgeev_args a) {
    if (batch <= 0 || a.n <= 0) return 0;
    a.ld_s = a.n | 1;
    a.use_smem = zgeev_smem_bytes(a.n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int threads = a.n <= 64 ? 128 : 256;
    return kh_launch<zgeev_args, zgeev_body>(dim3(batch), threads, zgeev_smem_bytes(a.n, a.ld_s, a.use_smem), st, a, "zgeev", 100.0 * a.n * a.n * a.n * batch);
}
