// Tiled variants of the eigensolver phases for matrices that do not fit in one CTA's shared memory
// (n > ~117: the 9x9 ... 15x15 harmonic bases, woodpile 11x11 and the extended-RCWA supercells).
//
//   zhb_* : BLOCKED Hessenberg reduction.  A panel kernel (one CTA per matrix) reduces nb columns while the
//           trailing matrix stays untouched in HBM/L2: step j needs p = A_j v and q^T = v^H A_j of the
//           partially updated matrix  A_j = A_0 - sum_i (x_i v_i^H + w_i q_i^T), which is ONE fused
//           read-only pass over A_0 (row dot products and column sums from the same loads) plus
//           corrections with the <= nb previous rank-2 terms.  The accumulated rank-2nb update
//               A[:, k:] -= [X W] [V^H; Q^T]
//           is then one DMMA GEMM (K = 2 nb) - 8 n^3 of the reduction's flops run on the tensor pipe and
//           the matrix is read once per COLUMN but written once per PANEL (the unblocked kernel reads and
//           writes it once per column).  Reflectors are kept as compact WY factors (V, T per panel); the
//           Schur basis Z = prod (I - V T V^H) is accumulated backwards, transposed, with two GEMMs per panel.
// Conventions as zgehd2 / zlarfg / zlarft: H_j = I - tau_j v_j v_j^H, A <- H_j^H A H_j, Z = H_0 H_1 ... H_{n-3}.
#pragma once
#include "kh_common.cuh"
#include "kh_zgemm.cuh"

struct zhb_args {
    int n, k0, nb;
    MatRef H;              // matrix in global memory, reduced in place
    MatRef XWt, VQ;        // [2 nb][n] per matrix: rows (x_i | w_i) and (conj v_i | q_i)
    MatRef Vc;             // [n][n]: column j = conj(v_j)                (B operand of the Z accumulation)
    MatRef VTt;            // [n][n]: rows k0 .. k0+nb-1 = (V T)^T of the panel
    cd* tau; long long tau_stride;
};

#define ZHB_NCH 8         /* column chunks (of one warp width) a lane accumulates per strip of the fused pass */
#define ZHB_QW 8          /* partial column-sum buffers in shared memory */

static inline size_t zhb_smem_bytes(int n, int nb) {
    return (size_t)(3 + ZHB_QW) * n * sizeof(cd) + (size_t)(5 * nb + nb * nb) * sizeof(cd) + 192 * sizeof(double) + 64;
}

KH_DEV void zhb_panel_body(const Cta& c, const zhb_args& a) {
    const int n = a.n, b = c.bx, k0 = a.k0, nb = a.nb;
    const int nbk = (nb < n - 2 - k0) ? nb : n - 2 - k0;                 // reflectors in this panel
    cd* H = mat_ptr(a.H, b);
    const int ld = a.H.ld;
    cd* XWt = mat_ptr(a.XWt, b);
    cd* VQ = mat_ptr(a.VQ, b);
    cd* Vc = mat_ptr(a.Vc, b);
    cd* VTt = mat_ptr(a.VTt, b);
    cd* tauout = a.tau + (long long)b * a.tau_stride;
    const int ldv = a.Vc.ld, ldt = a.VTt.ld;
    // shared: [v n][p n][q n][qpart QW x n][sa nb][sb nb][sc nb][sd nb][taus nb][T nb x nb][scratch]
    cd* vv = (cd*)KH_SMEM(c);
    cd* pv = vv + n;
    cd* qv = pv + n;
    cd* qpart = qv + n;
    cd* sa = qpart + ZHB_QW * n;
    cd* sb = sa + nb;
    cd* sc_ = sb + nb;
    cd* sd = sc_ + nb;
    cd* taus = sd + nb;
    cd* T = taus + nb;
    const int lane = c.tid % KH_WARP, warp = c.tid / KH_WARP, nw = (c.nthr + KH_WARP - 1) / KH_WARP;
#define XROW(r) (XWt + (long long)(r) * n)
#define VROW(r) (VQ + (long long)(r) * n)
    if (nbk < nb)                                                          // unused rows of the last panel: zero (K = 2 nb in the GEMM)
        for (int e = c.tid; e < 2 * nb * n; e += c.nthr) { XWt[e] = mk(0, 0); VQ[e] = mk(0, 0); }
    for (int e = c.tid; e < nb * nb; e += c.nthr) T[e] = mk(0, 0);
    c.sync();
    for (int jj = 0; jj < nbk; ++jj) {
        const int j = k0 + jj;
        // ---- P1: column j of the partially updated matrix
        for (int i = c.tid; i < n; i += c.nthr) {
            cd bv = H[(long long)i * ld + j];
            for (int m = 0; m < jj; ++m) {
                cfms(bv, XROW(m)[i], VROW(m)[j]);
                cfms(bv, XROW(nb + m)[i], VROW(nb + m)[j]);
            }
            vv[i] = bv;
        }
        c.sync();
        // ---- P2: Householder vector (every warp computes the norm redundantly)
        double part = 0.0;
        for (int i = j + 2 + lane; i < n; i += KH_WARP) part += cabs2(vv[i]);
        const double xn2 = kh_warp_allsum(part);
        const cd alpha = vv[j + 1];
        const bool trivial = (xn2 == 0.0 && alpha.y == 0.0);             // H_j = I (uniform across the CTA)
        cd tau = mk(0, 0), scl = mk(0, 0);
        if (!trivial) {
            const double beta = -copysign(sqrt(cabs2(alpha) + xn2), alpha.x);
            const double rbeta = 1.0 / beta;
            tau = mk((beta - alpha.x) * rbeta, -alpha.y * rbeta);
            scl = crecip(alpha - mk(beta, 0.0));
        }
        c.sync();                                                          // everyone has read the column
        for (int i = c.tid; i < n; i += c.nthr) {
            const cd vi = (i <= j) ? mk(0, 0) : ((i == j + 1) ? mk(1.0, 0.0) : vv[i] * scl);
            vv[i] = vi;
            VROW(jj)[i] = cconj(vi);
            Vc[(long long)i * ldv + j] = cconj(vi);
        }
        for (int w = warp; w < ZHB_QW; w += nw)
            for (int i = lane; i < n; i += KH_WARP) qpart[w * n + i] = mk(0, 0);
        for (int i = c.tid; i < n; i += c.nthr) pv[i] = mk(0, 0);
        if (c.tid == 0) { taus[jj] = tau; tauout[j] = tau; }
        c.sync();
        // ---- P3: scalars against the previous terms of the panel: a = v_m^H v, b = q_m^T v, c = v^H x_m  (one warp per dot)
        for (int job = warp; job < 3 * jj; job += nw) {
            const int m = job / 3, kind = job - 3 * m;
            const cd* row = (kind == 0) ? VROW(m) : ((kind == 1) ? VROW(nb + m) : XROW(m));
            cd acc = mk(0, 0);
            if (kind < 2) { for (int i = j + 1 + lane; i < n; i += KH_WARP) cfma(acc, row[i], vv[i]); }
            else { for (int i = j + 1 + lane; i < n; i += KH_WARP) cfma(acc, cconj(vv[i]), row[i]); }
            acc.x = kh_warp_allsum(acc.x); acc.y = kh_warp_allsum(acc.y);
            if (lane == 0) { if (kind == 0) sa[m] = acc; else if (kind == 1) sb[m] = acc; else sc_[m] = acc; }
        }
        // ---- P4: fused pass over A_0:  p0[i] = sum_c A[i][c] v[c],  q0[c] = sum_{i > j} conj(v[i]) A[i][c],  c >= j
        if (!trivial) {
            for (int cs = j; cs < n; cs += KH_WARP * ZHB_NCH) {              // strips of columns (column j included: q[j] = v^H b)
                cd qa[ZHB_NCH], vc[ZHB_NCH];
#pragma unroll
                for (int m = 0; m < ZHB_NCH; ++m) { qa[m] = mk(0, 0); const int cc = cs + lane + KH_WARP * m; vc[m] = cc < n ? vv[cc] : mk(0, 0); }
                for (int i = warp; i < n; i += nw) {
                    const cd* hr = H + (long long)i * ld;
                    cd hv[ZHB_NCH];
#pragma unroll
                    for (int m = 0; m < ZHB_NCH; ++m) { const int cc = cs + lane + KH_WARP * m; hv[m] = cc < n ? hr[cc] : mk(0, 0); }
                    cd pa = mk(0, 0);
                    const cd cvi = (i > j) ? cconj(vv[i]) : mk(0, 0);
#pragma unroll
                    for (int m = 0; m < ZHB_NCH; ++m) { cfma(pa, hv[m], vc[m]); cfma(qa[m], cvi, hv[m]); }
                    pa.x = kh_warp_allsum(pa.x); pa.y = kh_warp_allsum(pa.y);
                    if (lane == 0) pv[i] = pv[i] + pa;                     // (row i belongs to this warp in every strip)
                }
                // column sums: warps w and w + QW share a buffer, one after the other
                for (int round = 0; round * ZHB_QW < nw; ++round) {
                    if (warp / ZHB_QW == round) {
                        cd* qp = qpart + (warp % ZHB_QW) * n;
#pragma unroll
                        for (int m = 0; m < ZHB_NCH; ++m) { const int cc = cs + lane + KH_WARP * m; if (cc < n) qp[cc] = qp[cc] + qa[m]; }
                    }
                    c.sync();
                }
            }
        }
        c.sync();
        // ---- P5: corrections with the previous terms; q row out
        for (int i = c.tid; i < n; i += c.nthr) {
            cd p = pv[i], q = mk(0, 0);
            if (i >= j) for (int w = 0; w < ZHB_QW; ++w) q = q + qpart[w * n + i];
            for (int m = 0; m < jj; ++m) {
                const cd xm = XROW(m)[i], wm = XROW(nb + m)[i];
                cfms(p, xm, sa[m]); cfms(p, wm, sb[m]);
                if (i >= j) {
                    const cd dm = cconj(taus[m]) * cconj(sa[m]);          // v^H w_m
                    cfms(q, sc_[m], VROW(m)[i]); cfms(q, dm, VROW(nb + m)[i]);
                }
            }
            if (trivial) { p = mk(0, 0); q = mk(0, 0); }
            pv[i] = p; qv[i] = q;
            VROW(nb + jj)[i] = q;
        }
        c.sync();
        // ---- P6: s = v^H p (every warp redundantly), x = tau p - |tau|^2 s v, w = conj(tau) v, column jj of T
        double sr = 0.0, si = 0.0;
        for (int i = j + 1 + lane; i < n; i += KH_WARP) { const cd w = cconj(vv[i]) * pv[i]; sr += w.x; si += w.y; }
        const cd sv = mk(kh_warp_allsum(sr), kh_warp_allsum(si));
        const cd t2s = cabs2(tau) * sv, ctau = cconj(tau);
        for (int i = c.tid; i < n; i += c.nthr) {
            const cd vi = vv[i];
            XROW(jj)[i] = tau * pv[i] - t2s * vi;
            XROW(nb + jj)[i] = ctau * vi;
        }
        for (int r = c.tid; r <= jj; r += c.nthr) {
            cd t = mk(0, 0);
            if (r == jj) t = tau;
            else { for (int m = r; m < jj; ++m) cfma(t, T[r * nb + m], sa[m]); t = -(tau * t); }
            T[r * nb + jj] = t;
        }
        c.sync();
    }
    // ---- (V T)^T of the panel:  VTt[k0 + r][i] = sum_{m <= r} T[m][r] v_m[i]
    for (int e = c.tid; e < nbk * n; e += c.nthr) {
        const int r = e / n, i = e - r * n;
        cd acc = mk(0, 0);
        for (int m = 0; m <= r; ++m) cfma(acc, T[m * nb + r], cconj(VROW(m)[i]));
        VTt[(long long)(k0 + r) * ldt + i] = acc;
    }
#undef XROW
#undef VROW
}

// ---- balancing on a matrix in global memory (powers of two, EISPACK balanc criterion as the shared-memory kernels)
struct zbal_args { int n; MatRef A, H; cd* scale; long long scale_stride; cd* tau; long long tau_stride; };
KH_DEV void zbal_body(const Cta& c, const zbal_args& a) {
    const int n = a.n, b = c.bx;
    const cd* A = mat_ptr(a.A, b);
    cd* H = mat_ptr(a.H, b);
    const int ld = a.H.ld;
    cd* scout = a.scale + (long long)b * a.scale_stride;
    cd* tauout = a.tau + (long long)b * a.tau_stride;
    // shared: [rn n][cn n][f n][dsc n] doubles + scratch
    double* rn = (double*)KH_SMEM(c);
    double* cn = rn + n;
    double* fs = cn + n;
    double* dsc = fs + n;
    double* scratch = dsc + n;
    const int lane = c.tid % KH_WARP, warp = c.tid / KH_WARP, nw = (c.nthr + KH_WARP - 1) / KH_WARP;
    if (H != A)
        for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; H[(long long)i * ld + j] = A[(long long)i * a.A.ld + j]; }
    for (int i = c.tid; i < n; i += c.nthr) { dsc[i] = 1.0; tauout[i] = mk(0.0, 0.0); }
    c.sync();
    for (int sweep = 0; sweep < 12; ++sweep) {
        for (int i = warp; i < n; i += nw) {                               // row norms: one warp per row
            double s = 0.0;
            for (int j = lane; j < n; j += KH_WARP) if (j != i) s += cabs1(H[(long long)i * ld + j]);
            s = kh_warp_allsum(s);
            if (lane == 0) rn[i] = s;
        }
        for (int j = c.tid; j < n; j += c.nthr) {                          // column norms: one thread per column
            double s = 0.0;
            for (int i = 0; i < n; ++i) if (i != j) s += cabs1(H[(long long)i * ld + j]);
            cn[j] = s;
        }
        c.sync();
        double changed = 0.0;
        for (int i = c.tid; i < n; i += c.nthr) {
            const double cni = cn[i], rni = rn[i];
            double f = 1.0;
            if (cni != 0.0 && rni != 0.0 && cni <= 1e300 && rni <= 1e300) {
                double g = rni * 0.5, s = cni + rni, cc = cni;
                for (int q = 0; q < 1100 && cc < g; ++q) { f *= 2.0; cc *= 4.0; }
                g = rni * 2.0;
                for (int q = 0; q < 1100 && cc >= g; ++q) { f *= 0.5; cc *= 0.25; }
                if ((cc + rni) / f >= 0.95 * s) f = 1.0;
            }
            fs[i] = f;
            if (f != 1.0) changed = 1.0;
        }
        changed = cta_max(c, changed, scratch);
        c.sync();
        if (changed == 0.0) break;
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            double f = fs[j] / fs[i];
            if (f != 1.0) H[(long long)i * ld + j] = f * H[(long long)i * ld + j];
        }
        for (int i = c.tid; i < n; i += c.nthr) dsc[i] *= fs[i];
        c.sync();
    }
    for (int i = c.tid; i < n; i += c.nthr) scout[i] = mk(dsc[i], 0.0);
}

struct zhb_fin_args { int n; MatRef H, Zt; int mode; };      // mode 0: Zt = I ; mode 1: zero H below the sub-diagonal
KH_DEV void zhb_fin_body(const Cta& c, const zhb_fin_args& a) {
    const int n = a.n, b = c.bx;
    if (a.mode == 0) {
        cd* Zt = mat_ptr(a.Zt, b);
        for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; Zt[(long long)i * a.Zt.ld + j] = mk(i == j ? 1.0 : 0.0, 0.0); }
    } else {
        cd* H = mat_ptr(a.H, b);
        for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; if (j < i - 1) H[(long long)i * a.H.ld + j] = mk(0.0, 0.0); }
    }
}

static inline int zhb_nb(int n) { return n >= 384 ? 32 : 16; }
// work space (complex elements per matrix) of the blocked reduction besides H, Zt and the n x n buffer for conj(V)
static inline long long zhb_work_cd(int n) { return (long long)n * n + 4LL * zhb_nb(n) * n + 64; }

// A (balanced into H) -> Hessenberg H, Zt = Z^T.  Vc: n x n per matrix; work: zhb_work_cd(n) per matrix (stride wstride).
static inline int zhess_blocked_launch(kh_stream_t st, int batch, int n, MatRef A, MatRef H, MatRef Zt, MatRef Vc,
                                       cd* work, long long wstride, cd* scale, long long scale_stride, cd* tau, long long tau_stride, double workflops) {
    int e;
    {   zbal_args ba{n, A, H, scale, scale_stride, tau, tau_stride};
        if ((e = kh_launch<zbal_args, zbal_body>(dim3(batch), 512, (size_t)4 * n * sizeof(double) + 192 * sizeof(double) + 16, st, ba, "zgeev_hess", workflops))) return e; }
    if (n < 3) {
        zhb_fin_args f{n, H, Zt, 0};
        return kh_launch<zhb_fin_args, zhb_fin_body>(dim3(batch), 256, 0, st, f, "zgeev_hess", 0.0);
    }
    const int nb = zhb_nb(n);
    zhb_args a;
    a.n = n; a.nb = nb; a.H = H; a.Vc = Vc; a.tau = tau; a.tau_stride = tau_stride;
    a.VTt = mref(work, wstride, n);
    a.XWt = mref(work + (long long)n * n, wstride, n);
    a.VQ = mref(work + (long long)n * n + 2LL * nb * n, wstride, n);
    const size_t sm = zhb_smem_bytes(n, nb);
    for (int k0 = 0; k0 + 2 < n; k0 += nb) {
        a.k0 = k0;
        if ((e = kh_launch<zhb_args, zhb_panel_body>(dim3(batch), 512, sm, st, a, "zgeev_hess", 0.0))) return e;
        MatRef Hc = H; Hc.p = H.p + k0;
        MatRef Bq = a.VQ; Bq.p = a.VQ.p + k0;
        zgemm_args g = zgemm_make(n, n - k0, 2 * nb, a.XWt, Bq, Hc, -1.0);
        g.transA = 1; g.Cin = Hc; g.beta = 1.0;
        if ((e = zgemm_launch(st, batch, g))) return e;
    }
    {   zhb_fin_args f{n, H, Zt, 1};
        if ((e = kh_launch<zhb_fin_args, zhb_fin_body>(dim3(batch), 256, 0, st, f, "zgeev_hess", 0.0))) return e;
        f.mode = 0;
        if ((e = kh_launch<zhb_fin_args, zhb_fin_body>(dim3(batch), 256, 0, st, f, "zgeev_hess", 0.0))) return e; }
    // Z^T <- Z^T (I - conj(V) T^T V^T), panels in reverse order, on the trailing block that is not yet identity
    int klast = 0;
    while (klast + nb + 2 < n) klast += nb;
    MatRef W1 = a.XWt;                                                     // [m][nb] scratch (free now)
    for (int k0 = klast; k0 >= 0; k0 -= nb) {
        const int nbk = (nb < n - 2 - k0) ? nb : n - 2 - k0, m = n - k0 - 1;
        MatRef Zs = Zt; Zs.p = Zt.p + (long long)(k0 + 1) * Zt.ld + (k0 + 1);
        MatRef Vs = Vc; Vs.p = Vc.p + (long long)(k0 + 1) * Vc.ld + k0;
        MatRef Ts = a.VTt; Ts.p = a.VTt.p + (long long)k0 * a.VTt.ld + (k0 + 1);
        MatRef W1m = W1; W1m.ld = nb;
        zgemm_args g1 = zgemm_make(m, nbk, m, Zs, Vs, W1m);
        if ((e = zgemm_launch(st, batch, g1))) return e;
        zgemm_args g2 = zgemm_make(m, m, nbk, W1m, Ts, Zs, -1.0);
        g2.Cin = Zs; g2.beta = 1.0;
        if ((e = zgemm_launch(st, batch, g2))) return e;
    }
    return 0;
}
