// Batched non-Hermitian complex128 eigensolver, one CTA per matrix (replaces numpy.linalg.eig /
// LAPACK zgeev at khepri/alternative.py:172 for the Omega^2 = P Q problem of every patterned layer).
//
// Three kernels per batch (each phase gets its own launch shape and shows up separately in ncu):
//   zhessz: diagonal balancing with powers of two (exact similarity, as zgebal 'S') and Householder
//           reduction to upper Hessenberg form with the Schur vectors accumulated.  Z is kept
//           TRANSPOSED in HBM/L2 (Zt) so every column operation on Z is a coalesced row operation.
//   zqr   : single-shift QR iteration (Wilkinson / exceptional shifts, LAPACK zlahqr deflation test).
//           The Hessenberg matrix lives in shared memory in PACKED form (row i holds columns >= i-1),
//           79 KB at n = 98, so two CTAs share an SM.  Inside a sweep only the active window is
//           touched, with ONE barrier per Givens rotation: the column step of rotation k and the row
//           step of rotation k+1 touch disjoint entries except for a 2x2 corner that every thread
//           carries in registers.  The rotations of a sweep are stored and then applied in one pass
//           to the rows above / columns right of the window and to Z (coalesced, no barriers).
//   ztrevc: eigenvectors of the triangular factor by back substitution (as ztrevc), one thread per
//           eigenvector, lock-step so T is broadcast and X is accessed coalesced.
// The caller finishes with one DMMA GEMM:  W = diag(scale) * Z * X.
// Eigenvalue order / eigenvector scaling are free: parity is on physical outputs.
#pragma once
#include "kh_common.cuh"
#include "kh_zgeev_tiled.cuh"

struct zgeev_args {
    int n;
    MatRef A;          // input (only read by zhess; may alias Hw)
    MatRef Hw;         // n x n work matrix in global memory: Hessenberg form, then the triangular factor T
    MatRef Zt;         // out: transposed Schur vectors
    MatRef X;          // out: eigenvectors of T (upper triangular)
    cd* w; long long w_stride;          // out: eigenvalues
    cd* scale; long long scale_stride;  // out: balancing factors as complex (imag 0)
    cd* tau; long long tau_stride;      // work: Householder scalars of the Hessenberg reduction
    int* info;         // out: 0 ok, >0 = QR iteration failed to converge at that index+1
    int use_smem, ld_s;
    // Optional rotation log (work space, doubles).  When set, the QR kernel does not touch Z: it records every sweep
    // (window + rotations) and zrot_apply replays the log on Z afterwards, with Z resident in shared memory.
    // Per matrix: [0] header (int2: sweeps, rotations) [2 .. 2+sw_cap) sweep descriptors (int2: l | iact << 16, first
    // rotation) [2+sw_cap ..) rotations (c, re s, im s).
    double* rlog = nullptr; long long rlog_stride = 0; int rot_cap = 0, sw_cap = 0;
    // Phases of the QR iteration (only with a log).  With everything outside the active window deferred to the log, the
    // kernel only needs the leading na x na block: later phases run with a smaller shared-memory footprint and more
    // CTAs per SM.  A phase starts from istate[b] (phase > 0) and stops once iact < iact_stop.
    int na = 0, iact_stop = 0, phase = 0; int* istate = nullptr;
    // flag protocol of the relay sweep: a consumer that polls a progress counter more than spin_cap times gives up and the
    // matrix is reported as failed (info = n + 1).  stress (tests only): random pauses after every publication / before every poll.
    int spin_cap = 1 << 26, stress = 0;
};

#define KH_ENOMEM_ZGEEV (-2)
#define ZGEEV_EPS 2.220446049250313e-16   /* LAPACK ulp = eps*base */
#define ZGEEV_SAFMIN 2.2250738585072014e-308

struct kh_givens { double c; cd s, r; };
#ifndef KH_HOST_EMU
__device__ __forceinline__ cd kh_shfl_cd(cd v, int src) { return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
// release store to a shared-memory flag (publishes everything the warp wrote before the preceding __syncwarp)
__device__ __forceinline__ void kh_st_release_shared(int* p, int v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// explicit 32-bit shared-memory addressing for the sweep's inner loops (running addresses, immediate offsets)
__device__ __forceinline__ unsigned kh_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ cd kh_lds_cd(unsigned a) { cd v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ double kh_lds_d(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void kh_sts_cd(unsigned a, cd v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void kh_sts_d(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void kh_st_release_saddr(unsigned a, int v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// acquire load of a progress counter: everything the producer wrote before its release store is visible to the loads that follow
__device__ __forceinline__ int kh_ld_acquire_saddr(unsigned a) { int v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
// test switch (KH_QR_STRESS): pseudo-random pauses that shake the relative timing of producers and consumers
__device__ __forceinline__ void kh_stress_pause(int on, int salt) { if (on) __nanosleep((((unsigned)salt * 2654435761u) >> 23) & 0x1ff); }
#endif
// one entry of the rotation queue of a sweep (shared memory): G(t) = (c, s), and the corner entries after R(t)
struct alignas(16) kh_qrot { double c, pad; cd s, r, diag, nsub; };     // r = H[t][t-1], diag = H[t][t], nsub = H[t+1][t]
KH_DEV kh_givens make_givens(cd f, cd g) {
    // G = [[c, s], [-conj(s), c]],  G [f; g] = [r; 0];  c = |f|/h, s = (f/|f|) conj(g)/h, r = (f/|f|) h, h = sqrt(|f|^2+|g|^2).
    // With y = 1/sqrt(|f|^2 h^2):  c = |f|^2 y,  s = y f conj(g),  r = (h^2 y) f  -- ONE reciprocal square root, no
    // division, no branch (the degenerate cases are selects): the rotation sits on the critical path of the sweep.
    kh_givens G;
    const double g2 = cabs2(g), f2 = cabs2(f), h2 = f2 + g2;
    const bool f0 = (f2 == 0.0), g0 = (g2 == 0.0);
    const double y = kh_rsqrt(f0 ? g2 : f2 * h2);
    const cd fg = f * cconj(g);
    G.c = g0 ? 1.0 : (f0 ? 0.0 : f2 * y);
    G.s = g0 ? mk(0, 0) : (f0 ? y * cconj(g) : y * fg);
    G.r = g0 ? f : (f0 ? mk(g2 * y, 0.0) : (h2 * y) * f);
    return G;
}
#ifndef KH_HOST_EMU
// Same rotation for the generic case f != 0, g != 0 (the caller branches, warp-uniformly, to make_givens otherwise):
// no selects, and the reciprocal square root refined by ONE third-order step  y1 = y0 (1 + e/2 + 3e^2/8),
// e = 1 - x y0^2  (seed error 2^-20 -> 2^-58), four dependent operations instead of six.
__device__ __forceinline__ kh_givens make_givens_fast(cd f, cd g, double f2, double g2) {
    kh_givens G;
    const double h2 = f2 + g2, x = f2 * h2;
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-(x * y0), y0, 1.0);
    const double y = fma(y0 * e, fma(0.375, e, 0.5), y0);
    const cd fg = f * cconj(g);
    G.c = f2 * y; G.s = y * fg; G.r = (h2 * y) * f;
    return G;
}
#endif

#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
__device__ long long kh_qr_dbg[16];
#define QT_DECL long long qt0 = clock64(), qt_scan = 0, qt_shift = 0, qt_sweep = 0, qt_delay = 0, qt_n = 0, qt_rot = 0, qt_t
#define QT_MARK() (qt_t = clock64())
#define QT_ADD(v) do { long long _n = clock64(); (v) += _n - qt_t; qt_t = _n; } while (0)
#else
#define QT_DECL
#define QT_MARK()
#define QT_ADD(v)
#endif
// ============================================================================ 1a. fused shared-memory variant
// Balance + Hessenberg reduction + formation of Z in ONE kernel with the matrix resident in shared memory (n <= ~117).
// Per reduction step three barriers: [norm, v] | [p = A v, q = v^H A with several threads per dot product, s = v^H p] |
// [rank-2 update, thread <-> row so that its coefficients stay in registers, four columns in flight].  Z is then
// accumulated IN PLACE (LAPACK zunghr/zung2r layout: reflector k sits in column k+1 until step k consumes it), two
// barriers per step and no global-memory access inside the loop.
KH_DEV void zhessz_body(const Cta& c, const zgeev_args& a) {
    const int n = a.n, b = c.bx, ld = a.ld_s;
    const cd* A = mat_ptr(a.A, b);
    cd* Hg = mat_ptr(a.Hw, b);
    cd* Ztg = mat_ptr(a.Zt, b);
    const int ldz = a.Zt.ld, ldg = a.Hw.ld;
    cd* scout = a.scale + (long long)b * a.scale_stride;
    // shared: [vv n][pp n][scratch 192 dbl][qq 2n cd (balancing factors first)][H n x ld]
    cd* vv = (cd*)KH_SMEM(c);
    cd* pp = vv + n;
    double* scratch = (double*)(pp + n);
    double* dsc = scratch + 192;
    cd* vq = (cd*)dsc;                  // [n][2]: (conj v_j, q_j) interleaved for the update
    cd* Hs = (cd*)(KH_SMEM(c) + (((2 * n * 16 + 192 * 8 + 4 * n * 8) + 15) & ~15));
    cd* spart = (cd*)scratch;           // [<= 32] per-warp partial sums of s = v^H p
    cd* tauout = a.tau + (long long)b * a.tau_stride;
#define HH(i, j) Hs[(i) * ld + (j)]
    for (int i = c.tid; i < n; i += c.nthr) { dsc[i] = 1.0; tauout[i] = mk(0.0, 0.0); }
    kh_stage_rows(c, Hs, ld, A, a.A.ld, n, n, (unsigned long long*)(scratch + 190));     // the matrix comes in by TMA bulk copies, one row each
    // ---- balancing (Jacobi-style sweeps of the EISPACK balanc criterion; powers of two, so exact)
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
            pp[i].x = f;
            if (f != 1.0) changed = 1.0;
        }
        changed = cta_max(c, changed, scratch);
        c.sync();
        if (changed == 0.0) break;
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            double f = pp[j].x / pp[i].x;
            if (f != 1.0) HH(i, j) = f * HH(i, j);
        }
        for (int i = c.tid; i < n; i += c.nthr) dsc[i] *= pp[i].x;
        c.sync();
    }
    for (int i = c.tid; i < n; i += c.nthr) scout[i] = mk(dsc[i], 0.0);
    c.sync();

    const int lane = c.tid % KH_WARP, warp = c.tid / KH_WARP, nw = (c.nthr + KH_WARP - 1) / KH_WARP;
#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
    long long ht = clock64(), h_norm = 0, h_mv = 0, h_upd = 0, h_out = 0, h_zb = 0, h_zc = 0, h_n;
#define HT_ADD(v) do { h_n = clock64(); (v) += h_n - ht; ht = h_n; } while (0)
#else
#define HT_ADD(v)
#endif
    // update phase: RG row groups of RPG rows (thread <-> row), CB interleaved column classes
    const int RG = (n + 31) / 32, RPG = (n + RG - 1) / RG;
#ifdef KH_HOST_EMU
    const int CB = 1, VT = RG * 32;
#else
    const int CB = (nw / RG) > 0 ? (nw / RG) : 1, VT = RG * CB * 32;
#endif
    for (int k = 0; k + 2 < n; ++k) {
        // ---- phase 1: norm of the column below the sub-diagonal (every warp redundantly), Householder scalars, v
        double part = 0.0;
        for (int i = k + 2 + lane; i < n; i += KH_WARP) part += cabs2(HH(i, k));
        const double xn2 = kh_warp_allsum(part);
        const cd alpha = HH(k + 1, k);
        if (xn2 == 0.0 && alpha.y == 0.0) continue;          // already reduced: H_k = I   (uniform across the CTA)
        // sqrt, 1/beta and 1/(alpha - beta) sit on the critical path of every step (all threads wait for them): one reciprocal
        // square root gives both |beta| and 1/beta, and the complex reciprocal is division free (kh_crecip_fast); the library
        // sqrt + two divisions + Smith's complex division were ~2 k cycles of dependent FP64 work per Householder step
        const double s2 = cabs2(alpha) + xn2, rs = kh_rsqrt(s2);
        const double beta = -copysign(s2 * rs, alpha.x);
        const double rbeta = -copysign(rs, alpha.x);
        const cd tau = mk((beta - alpha.x) * rbeta, -alpha.y * rbeta);
#ifdef KH_HOST_EMU
        const cd sc = crecip(alpha - mk(beta, 0.0));
#else
        const cd sc = kh_crecip_fast(alpha - mk(beta, 0.0));
#endif
        const cd ctau = cconj(tau);
        for (int i = k + 1 + c.tid; i < n; i += c.nthr) {
            const cd vi = (i == k + 1) ? mk(1.0, 0.0) : HH(i, k) * sc;
            vv[i] = vi;
            vq[2 * i] = cconj(vi);
        }
        if (c.tid == 0) tauout[k] = tau;
        c.sync();
        HT_ADD(h_norm);
        // ---- phase 2: p = A v (jobs 0..n-1, one per row), q = v^H A (jobs n.., one per column > k), TPJ threads per job
        const int m = n - k - 1, J = n + m;
#ifdef KH_HOST_EMU
        const int TPJ = 1;
#else
        const int TPJ = (4 * J <= c.nthr) ? 4 : ((2 * J <= c.nthr) ? 2 : 1);
#endif
        cd contrib = mk(0, 0);
        const int jpp = c.nthr / TPJ, sub = c.tid % TPJ;
        for (int jb = 0; jb < J; jb += jpp) {
            const int job = jb + c.tid / TPJ;
            cd a0 = mk(0, 0), a1 = mk(0, 0);
            if (job < n) {
                const cd* hr = &HH(job, 0);
                int j = k + 1 + sub;
                for (; j + TPJ < n; j += 2 * TPJ) { cfma(a0, hr[j], vv[j]); cfma(a1, hr[j + TPJ], vv[j + TPJ]); }
                if (j < n) cfma(a0, hr[j], vv[j]);
            } else if (job < J) {
                const int jc = job - n + k + 1;
                const cd* hc = &HH(0, jc);
                int i = k + 1 + sub;
                for (; i + TPJ < n; i += 2 * TPJ) { cfma(a0, vq[2 * i], hc[i * ld]); cfma(a1, vq[2 * (i + TPJ)], hc[(i + TPJ) * ld]); }
                if (i < n) cfma(a0, vq[2 * i], hc[i * ld]);
            }
            cd acc = a0 + a1;
#ifndef KH_HOST_EMU
            for (int o = TPJ >> 1; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
#endif
            if (sub == 0) {
                if (job < n) { pp[job] = acc; if (job > k) contrib = contrib + vq[2 * job] * acc; }
                else if (job < J) vq[2 * (job - n + k + 1) + 1] = acc;
            }
        }
        contrib.x = kh_warp_allsum(contrib.x); contrib.y = kh_warp_allsum(contrib.y);
        if (lane == 0) spart[warp] = contrib;
        c.sync();
        HT_ADD(h_mv);
        // ---- phase 3: A -= pt conj(v)^T + (conj(tau) v) q^T,  pt = tau p - |tau|^2 s v ; the reflector goes into column k
        cd sv = mk(0, 0);
        for (int w = 0; w < nw; ++w) sv = sv + spart[w];
        const cd t2s = cabs2(tau) * sv;
        for (int vt = c.tid; vt < VT; vt += c.nthr) {
            const int vw = vt >> 5, vl = vt & 31, rg = vw % RG, cb = vw / RG;
            const int i = rg * RPG + vl;
            if (vl >= RPG || i >= n) continue;
            cd pt = tau * pp[i], tv = mk(0, 0);
            if (i > k) { const cd vi = vv[i]; tv = ctau * vi; pt = pt - t2s * vi; }
            cd* hr = &HH(i, 0);
            int j = k + 1 + cb;
            for (; j + 3 * CB < n; j += 4 * CB) {
                cd h0 = hr[j], h1 = hr[j + CB], h2 = hr[j + 2 * CB], h3 = hr[j + 3 * CB];
                cfms(h0, pt, vq[2 * j]); cfms(h1, pt, vq[2 * (j + CB)]); cfms(h2, pt, vq[2 * (j + 2 * CB)]); cfms(h3, pt, vq[2 * (j + 3 * CB)]);
                cfms(h0, tv, vq[2 * j + 1]); cfms(h1, tv, vq[2 * (j + CB) + 1]); cfms(h2, tv, vq[2 * (j + 2 * CB) + 1]); cfms(h3, tv, vq[2 * (j + 3 * CB) + 1]);
                hr[j] = h0; hr[j + CB] = h1; hr[j + 2 * CB] = h2; hr[j + 3 * CB] = h3;
            }
            for (; j < n; j += CB) { cd h = hr[j]; cfms(h, pt, vq[2 * j]); cfms(h, tv, vq[2 * j + 1]); hr[j] = h; }
            if (cb == 0 && i > k) hr[k] = (i == k + 1) ? mk(beta, 0.0) : vv[i];       // reflector kept in place
        }
        c.sync();
        HT_ADD(h_upd);
    }
    // ---- Hessenberg matrix out (the QR kernel reads j >= i-1 only)
    for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; Hg[(long long)i * ldg + j] = (j >= i - 1) ? HH(i, j) : mk(0.0, 0.0); }
    c.sync();
    // ---- Z = H_0 H_1 ... H_{n-3} in place: shift the reflectors one column to the right, identity elsewhere
    for (int i = c.tid; i < n; i += c.nthr) {
        for (int j = n - 1; j >= 0; --j)
            HH(i, j) = (j >= 1 && j <= i - 1) ? HH(i, j - 1) : mk(i == j ? 1.0 : 0.0, 0.0);
        vq[i] = tauout[i];                                                  // (written by this CTA; vq is free now)
    }
    c.sync();
    HT_ADD(h_out);
    for (int k = n - 3; k >= 0; --k) {
        const cd tau = vq[k];
        if (tau.x == 0.0 && tau.y == 0.0) continue;                       // uniform
        // w_j = v^H Z[k+1:, j] for j >= k+2 (v = [1; column k+1 below the diagonal]); Z[k+1][j] is still zero there
        const int m = n - k - 2;
#ifdef KH_HOST_EMU
        const int TPJ = 1;
#else
        const int TPJ = (4 * m <= c.nthr) ? 4 : ((2 * m <= c.nthr) ? 2 : 1);
#endif
        const int jpp = c.nthr / TPJ, sub = c.tid % TPJ;
        for (int i = k + 1 + c.tid; i < n; i += c.nthr) vv[i] = (i == k + 1) ? mk(1.0, 0.0) : HH(i, k + 1);
        for (int jb = 0; jb < m; jb += jpp) {
            const int job = jb + c.tid / TPJ;
            cd a0 = mk(0, 0), a1 = mk(0, 0);
            if (job < m) {
                const cd* zc = &HH(0, k + 2 + job);
                const cd* vc = &HH(0, k + 1);
                int i = k + 2 + sub;
                for (; i + TPJ < n; i += 2 * TPJ) { cfma(a0, cconj(vc[i * ld]), zc[i * ld]); cfma(a1, cconj(vc[(i + TPJ) * ld]), zc[(i + TPJ) * ld]); }
                if (i < n) cfma(a0, cconj(vc[i * ld]), zc[i * ld]);
            }
            cd acc = a0 + a1;
#ifndef KH_HOST_EMU
            for (int o = TPJ >> 1; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
#endif
            if (sub == 0 && job < m) pp[k + 2 + job] = tau * acc;
        }
        c.sync();
        HT_ADD(h_zb);
        // Z[i][j] -= v_i (tau w_j) for i >= k+1, j >= k+2 ; column k+1 becomes e_{k+1} - tau v
        for (int vt = c.tid; vt < VT; vt += c.nthr) {
            const int vw = vt >> 5, vl = vt & 31, rg = vw % RG, cb = vw / RG;
            const int i = rg * RPG + vl;
            if (vl >= RPG || i >= n || i <= k) continue;
            const cd vi = vv[i];
            cd* zr = &HH(i, 0);
            int j = k + 2 + cb;
            for (; j + 3 * CB < n; j += 4 * CB) {
                cd z0 = zr[j], z1 = zr[j + CB], z2 = zr[j + 2 * CB], z3 = zr[j + 3 * CB];
                cfms(z0, vi, pp[j]); cfms(z1, vi, pp[j + CB]); cfms(z2, vi, pp[j + 2 * CB]); cfms(z3, vi, pp[j + 3 * CB]);
                zr[j] = z0; zr[j + CB] = z1; zr[j + 2 * CB] = z2; zr[j + 3 * CB] = z3;
            }
            for (; j < n; j += CB) { cd z = zr[j]; cfms(z, vi, pp[j]); zr[j] = z; }
            if (cb == 0) zr[k + 1] = (i == k + 1) ? mk(1.0 - tau.x, -tau.y) : mk(0, 0) - tau * vi;
        }
        c.sync();
        HT_ADD(h_zc);
    }
#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
    if (c.tid == 0 && b == 0) { kh_qr_dbg[8] += h_norm; kh_qr_dbg[9] += h_mv; kh_qr_dbg[10] += h_upd; kh_qr_dbg[11] += h_out; kh_qr_dbg[12] += h_zb; kh_qr_dbg[13] += h_zc; }
#endif
    for (int e = c.tid; e < n * n; e += c.nthr) { int j = e / n, i = e - j * n; Ztg[(long long)j * ldz + i] = HH(i, j); }
#undef HH
}

// ============================================================================ 2. shifted QR on the packed Hessenberg matrix
// packed row-major upper Hessenberg: row i holds columns max(i-1,0) .. n-1; HQ(i,j) = Hp[hp_off(i) + j]
KH_HD int hp_off(int i, int n) { return i * n - ((i - 1) * i) / 2; }
KH_HD int hp_size(int n) { return hp_off(n - 1, n) + n + 3; }

// ---- tiled sweep for a Hessenberg matrix in global memory (n beyond shared memory).  The rotations of a sweep are
// generated in groups of ZQT_W.  The diagonal tile that determines a group (rows ks..ke, columns ks-1..ke) is staged in
// shared memory and chased there by ONE warp, lane <-> tile column: the dependent chain
//     corner (shuffle) -> C(t-1) on rows t, t+1 in registers -> Givens G(t) -> R(t) on the own column
// runs without any barrier or shared-memory round trip (each lane keeps the running bottom entry of its column in a
// register, exactly like the relay sweep of the packed kernel); the column steps C(t) on the tile rows above the corner
// are applied afterwards, lane <-> row, from the queue.  The rest of the window then receives the whole group in bulk:
// one thread per column right of the tile (row rotations, rows ks..ke) and one thread per row above it (column
// rotations, ALL rows 0..ks-1, so nothing above the window is left for later) - each element is loaded and stored once
// per group instead of once per rotation, with eight loads in flight per thread.
// State between groups: R(t), t < ke, applied everywhere; C(ke-1) still pending on rows ke, ke+1 (the next group's first
// chain step applies it), H[ke][ke-1] holds the sub-diagonal entry after R(ke-1).
#define ZQT_W 28
#define ZQT_LD 33
template <bool ROWS>      // ROWS: row rotations down a column (stride = ld); else column rotations along a row (stride 1)
KH_DEV void zqt_bulk_chain(cd* hp, long long stride, const kh_qrot* Q, int ks, int ke) {
    cd h0 = hp[0];
    int k = ks;
    for (; k + 8 <= ke; k += 8, hp += 8 * stride) {
        cd h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) h[u] = hp[(u + 1) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double cc = Q[k + u].c; const cd ss = Q[k + u].s;
            if (ROWS) { hp[u * stride] = cc * h0 + ss * h[u]; h0 = cc * h[u] - cconj(ss) * h0; }
            else { hp[u * stride] = cc * h0 + cconj(ss) * h[u]; h0 = cc * h[u] - ss * h0; }
        }
    }
    for (; k < ke; ++k, hp += stride) {
        const cd h1 = hp[stride];
        const double cc = Q[k].c; const cd ss = Q[k].s;
        if (ROWS) { hp[0] = cc * h0 + ss * h1; h0 = cc * h1 - cconj(ss) * h0; }
        else { hp[0] = cc * h0 + cconj(ss) * h1; h0 = cc * h1 - ss * h0; }
    }
    hp[0] = h0;
}
KH_DEV void zqr_tiled_sweep(const Cta& c, cd* Hg, int ldg, int l, int iact, cd f_first, cd g_first, kh_qrot* Q, cd* D) {
    const int lane = c.tid % KH_WARP, warp = c.tid / KH_WARP;
#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
    long long tt = clock64(), tn, t_load = 0, t_chase = 0, t_bulk = 0;
#define TT_ADD(v) do { tn = clock64(); (v) += tn - tt; tt = tn; } while (0)
#else
#define TT_ADD(v)
#endif
#ifdef KH_HOST_EMU
    // ---- host emulation (one virtual thread): same grouping, every group completes its column steps on rows <= k+2
    for (int ks = l; ks < iact; ks += ZQT_W) {
        const int ke = (ks + ZQT_W < iact) ? ks + ZQT_W : iact;
        const int rA = ks, rB = (ke + 1 < iact) ? ke + 1 : iact;
        const int cA = (ks - 1 > l) ? ks - 1 : l, cB = rB;
        const int nr = rB - rA + 1, nc = cB - cA + 1;
        for (int e = c.tid; e < nr * nc; e += c.nthr) { const int i = e / nc, j = e - i * nc; D[i * (ZQT_LD + 2) + j] = Hg[(long long)(rA + i) * ldg + cA + j]; }
#define DT(i, j) D[((i) - rA) * (ZQT_LD + 2) + ((j) - cA)]
        for (int k = ks; k < ke; ++k) {
            kh_givens G;
            if (k == l) G = make_givens(f_first, g_first);
            else { G = make_givens(DT(k, k - 1), DT(k + 1, k - 1)); DT(k, k - 1) = G.r; DT(k + 1, k - 1) = mk(0.0, 0.0); }
            const cd cs = cconj(G.s);
            for (int j = k; j <= cB; ++j) { const cd h0 = DT(k, j), h1 = DT(k + 1, j); DT(k, j) = G.c * h0 + G.s * h1; DT(k + 1, j) = G.c * h1 - cs * h0; }
            const int rl = (k + 2 < rB) ? k + 2 : rB;
            for (int r = rA; r <= rl; ++r) { const cd h0 = DT(r, k), h1 = DT(r, k + 1); DT(r, k) = G.c * h0 + cs * h1; DT(r, k + 1) = G.c * h1 - G.s * h0; }
            Q[k].c = G.c; Q[k].s = G.s;
        }
#undef DT
        for (int e = c.tid; e < nr * nc; e += c.nthr) { const int i = e / nc, j = e - i * nc; Hg[(long long)(rA + i) * ldg + cA + j] = D[i * (ZQT_LD + 2) + j]; }
        for (int t = 0; t < iact - cB; ++t) zqt_bulk_chain<true>(Hg + (long long)ks * ldg + (cB + 1 + t), ldg, Q, ks, ke);
        for (int t = 0; t < rA; ++t) zqt_bulk_chain<false>(Hg + (long long)t * ldg + ks, 1, Q, ks, ke);
    }
#else
    for (int ks = l; ks < iact; ks += ZQT_W) {
        const int ke = (ks + ZQT_W < iact) ? ks + ZQT_W : iact;          // rotations ks .. ke-1
        const int rA = ks, cA = (ks > l) ? ks - 1 : l;                    // tile rows ks..ke, columns cA..ke
        const int nr = ke - rA + 1, nc = ke - cA + 1;
        for (int e = c.tid; e < nr * nc; e += c.nthr) { const int i = e / nc, j = e - i * nc; D[i * ZQT_LD + j] = Hg[(long long)(rA + i) * ldg + cA + j]; }
        __syncthreads();
        TT_ADD(t_load);
        if (warp == 0) {
#define DT(i, j) D[((i) - rA) * ZQT_LD + ((j) - cA)]
            const int j = cA + lane, jc = (j <= ke) ? j : ke;             // own tile column (clamped for the unconditional loads)
            const bool own = (j <= ke);
            kh_givens G;
            cd sub, carry;
            int t;
            if (ks == l) {                                                // first group: G(l) from the shift, R(l) on every column
                G = make_givens(f_first, g_first);
                const cd h0 = DT(l, jc), h1 = DT(l + 1, jc);
                carry = G.c * h1 - cconj(G.s) * h0;
                if (own) DT(l, j) = G.c * h0 + G.s * h1;
                sub = kh_shfl_cd(carry, 0);
                if (lane == 0) { Q[l].c = G.c; Q[l].s = G.s; }
                t = l + 1;
            } else {                                                      // continue: G(ks-1) is in the queue
                G.c = Q[ks - 1].c; G.s = Q[ks - 1].s; G.r = mk(0, 0);
                carry = DT(ks, jc);
                sub = DT(ks, ks - 1);
                t = ks;
            }
            for (; t < ke; ++t) {
                const cd hd = DT(t + 1, t);                                // untouched by this sweep so far
                const cd h1 = DT(t + 1, jc);
                const cd cb = kh_shfl_cd(carry, t - cA);                   // H[t][t] after R(t-1)
                const cd a1 = G.c * sub + cconj(G.s) * cb, b1 = G.c * cb - G.s * sub;      // C(t-1) on row t
                const cd c1 = cconj(G.s) * hd, d1 = G.c * hd;                              // C(t-1) on row t+1: bulge
                const double f2 = cabs2(a1), g2 = cabs2(c1);
                kh_givens Gn;
                if (f2 == 0.0 || g2 == 0.0) Gn = make_givens(a1, c1);
                else Gn = make_givens_fast(a1, c1, f2, g2);
                const cd ncs = cconj(Gn.s);
                const cd top = Gn.c * carry + Gn.s * h1, bot = Gn.c * h1 - ncs * carry;    // R(t) on the own column
                const cd newdiag = Gn.c * b1 + Gn.s * d1, nextsub = Gn.c * d1 - ncs * b1;
                if (own && j > t) { DT(t, j) = top; carry = bot; }
                if (lane == 0) { Q[t].c = Gn.c; Q[t].s = Gn.s; Q[t].r = Gn.r; Q[t].diag = newdiag; Q[t].nsub = nextsub; }
                G = Gn; sub = nextsub;
            }
            {   const cd cb = kh_shfl_cd(carry, ke - cA);                  // row ke of column ke after R(ke-1)
                __syncwarp();
                if (lane == 0) {
                    if (ke == iact) { DT(ke, ke - 1) = G.c * sub + cconj(G.s) * cb; DT(ke, ke) = G.c * cb - G.s * sub; }   // C(iact-1) on row iact
                    else { DT(ke, ke - 1) = sub; DT(ke, ke) = cb; }
                }
            }
            __syncwarp();
            // column steps C(t), t >= r, on the tile rows r < ke (rows t+1, t+2 were handled inside the chain)
            const int r = rA + lane;
            if (r < ke) {
                cd car = mk(0, 0);
                for (int t2 = r; t2 < ke; ++t2) {
                    const double cc = Q[t2].c; const cd ss = Q[t2].s;
                    cd h0 = car;
                    if (t2 == r) { if (t2 > l) { h0 = Q[t2].diag; DT(r, t2 - 1) = Q[t2].r; } else h0 = DT(r, t2); }
                    const cd h1 = DT(r, t2 + 1);
                    DT(r, t2) = cc * h0 + cconj(ss) * h1;
                    car = cc * h1 - ss * h0;
                }
                DT(r, ke) = car;
            }
#undef DT
        }
        __syncthreads();
        TT_ADD(t_chase);
        for (int e = c.tid; e < nr * nc; e += c.nthr) { const int i = e / nc, j = e - i * nc; Hg[(long long)(rA + i) * ldg + cA + j] = D[i * ZQT_LD + j]; }
        const int nRight = iact - ke, nUp = rA;
        for (int t = c.tid; t < nRight + nUp; t += c.nthr) {
            if (t < nRight) zqt_bulk_chain<true>(Hg + (long long)ks * ldg + (ke + 1 + t), ldg, Q, ks, ke);
            else zqt_bulk_chain<false>(Hg + (long long)(t - nRight) * ldg + ks, 1, Q, ks, ke);
        }
        __syncthreads();
        TT_ADD(t_bulk);
    }
#if defined(KH_QR_TIMING)
    if (c.tid == 0 && c.bx == 0) { kh_qr_dbg[8] += t_load; kh_qr_dbg[9] += t_chase; kh_qr_dbg[10] += t_bulk; }
#endif
#endif
}

template <bool PACKED>
KH_DEV void zqr_body_t(const Cta& c, const zgeev_args& a) {
    const int nfull = a.n, n = (a.na > 0 && a.na < a.n) ? a.na : a.n, b = c.bx;     // n: leading block this phase works on
    int iact = n - 1;
    if (a.phase > 0) {                                                // continue where the previous phase stopped
        iact = a.istate[b];
        if (iact < a.iact_stop || iact < 0) return;                   // nothing left for this phase (uniform)
        if (iact > n - 1) iact = n - 1;
    }
    cd* Hg = mat_ptr(a.Hw, b);
    cd* Zt = mat_ptr(a.Zt, b);
    const int ldz = a.Zt.ld, ldg = a.Hw.ld;
    cd* wout = a.w + (long long)b * a.w_stride;
    // shared: [Q: n rotation-queue entries][ctl 16 int][packed H]  (H stays in global memory, full storage, when it does not fit)
    kh_qrot* Q = (kh_qrot*)KH_SMEM(c); // rotation queue of the current sweep
    int* ctl = (int*)(Q + n);          // [0] deflation scan, [2] error flag, [4..7] per-column-warp progress counters
    cd* Hp = (cd*)(KH_SMEM(c) + (((n * (int)sizeof(kh_qrot) + 16 * 4) + 15) & ~15));   // offset arithmetic keeps the shared address space
    const bool packed = PACKED;
    cd* const Hb = PACKED ? Hp : Hg;
#define ROWOFF(i) (PACKED ? hp_off((i), n) : (i) * ldg)
#define ROWSTEP(i) (PACKED ? (n - (i)) : ldg)          /* ROWOFF(i+1) - ROWOFF(i) */
#define HQ(i, j) Hb[ROWOFF(i) + (j)]
#define ZT(i, j) Zt[(long long)(i) * ldz + (j)]
    if (packed)
        for (int e = c.tid; e < n * n; e += c.nthr) { int i = e / n, j = e - i * n; if (j >= i - 1) HQ(i, j) = Hg[(long long)i * ldg + j]; }
    if (c.tid == 0) ctl[2] = 0;
    c.sync();

    const double smlnum = ZGEEV_SAFMIN * ((double)nfull / ZGEEV_EPS);
    const int itmax = 30 * (nfull > 10 ? nfull : 10);
    int fail = 0;
    int its = 0;
    double* const rl = a.rlog ? a.rlog + (long long)b * a.rlog_stride : nullptr;
    int* const rl_sw = rl ? (int*)(rl + 2) : nullptr;
    double* const rl_rot = rl ? rl + 2 + a.sw_cap : nullptr;
    int nsw = 0, nrot = 0;             // logged sweeps / rotations (uniform)
    if (rl && a.phase > 0) { nsw = ((const int*)rl)[0]; nrot = ((const int*)rl)[1]; }
    const int istop = a.iact_stop > 0 ? a.iact_stop : 0;
    QT_DECL;
    while (iact >= istop) {
        QT_MARK();
        // ---- locate the active block [l, iact] (zlahqr deflation criterion)
        if (c.tid == 0) ctl[0] = 0;
        c.sync();
        for (int k = iact - c.tid; k >= 1; k -= c.nthr) {
            cd hs = HQ(k, k - 1);
            bool negl = false;
            if (cabs1(hs) <= smlnum) negl = true;
            else {
                double tst = cabs1(HQ(k - 1, k - 1)) + cabs1(HQ(k, k));
                if (tst == 0.0) {
                    if (k - 2 >= 0) tst += cabs1(HQ(k - 1, k - 2));
                    if (k + 1 <= n - 1) tst += cabs1(HQ(k + 1, k));
                }
                if (cabs1(hs) <= ZGEEV_EPS * tst) {
                    double h12 = cabs1(HQ(k - 1, k)), h21 = cabs1(hs);
                    double ab = fmax(h21, h12), ba = fmin(h21, h12);
                    double d1 = cabs1(HQ(k, k)), d2 = cabs1(HQ(k - 1, k - 1) - HQ(k, k));
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
        if (l > 0 && c.tid == 0) HQ(l, l - 1) = mk(0.0, 0.0);
        if (l >= iact) {                       // one eigenvalue converged
            if (c.tid == 0) wout[iact] = HQ(iact, iact);
            iact -= 1; its = 0;
            c.sync();
            continue;
        }
        its += 1;
        if (its > itmax) { fail = iact + 1; break; }
        QT_ADD(qt_scan);
        // ---- shift (zlahqr)
        cd t;
        if (its % 10 == 0 && (its / 10) % 2 == 1) t = HQ(l, l) + mk(0.75 * cabs1(HQ(l + 1, l)), 0.0);
        else if (its % 10 == 0) t = HQ(iact, iact) + mk(0.75 * cabs1(HQ(iact, iact - 1)), 0.0);
        else {
            t = HQ(iact, iact);
#ifdef KH_HOST_EMU
            cd u = csqrt_(HQ(iact - 1, iact)) * csqrt_(HQ(iact, iact - 1));
            double s = cabs1(u);
            if (s != 0.0) {
                cd x = 0.5 * (HQ(iact - 1, iact - 1) - t);
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
#else
            // Wilkinson shift with division-free scalar arithmetic: every thread of the CTA evaluates it before the sweep can
            // start, and the three complex square roots + three divisions of the textbook form were several thousand cycles of
            // dependent FP64 work per sweep.  (A shift only steers convergence: a few ulp in it do not touch the similarity.)
            cd u = kh_csqrt_fast(HQ(iact - 1, iact)) * kh_csqrt_fast(HQ(iact, iact - 1));
            double s = cabs1(u);
            if (s != 0.0) {
                cd x = 0.5 * (HQ(iact - 1, iact - 1) - t);
                const double sx = cabs1(x);
                s = fmax(s, sx);
                const double rs = kh_rcp_fast(s);
                cd xs = rs * x, us = rs * u;
                cd y = s * kh_csqrt_fast(xs * xs + us * us);
                if (sx > 0.0 && x.x * y.x + x.y * y.y < 0.0) y = -y;
                t = t - u * (u * kh_crecip_fast(x + y));
            }
#endif
        }
        // ---- one implicit single-shift sweep over the window [l, iact], one barrier per rotation.
        // R(k): rows k,k+1 <- G_k (window columns);  C(k): columns k,k+1 <- . G_k^H (window rows <= k+2).
        // Thread t owns index m = l + t: it applies the row steps to column m+1 while k < m and the column
        // steps to row m once k >= m, so its addresses advance by running offsets (no multiplies in the loop).
        const cd f_first = HQ(l, l) - t, g_first = HQ(l + 1, l);
        QT_ADD(qt_shift);
#ifndef KH_HOST_EMU
        if (PACKED && c.nthr >= 64 && n <= 32 * (c.nthr >> 6)) {
            // ===== warp-specialised RELAY sweep (GPU, packed path).  The first half of the warps own the window columns
            // (thread <-> column l + 32 w + lane, the running bottom entry of the column stays in a register), the second
            // half own the window rows.  Rotation t is GENERATED by the warp that owns column t: it has the corner
            // H[t][t] in a lane, runs the dependent chain (corner -> Givens -> own column -> shuffle) without any block
            // barrier, and publishes (c, s, r, diagonal, sub-diagonal) in a shared-memory queue.  Column warps to the right
            // FOLLOW (apply the published rotations to their columns) until the corner reaches them and they take over
            // the chain; row warps apply the column steps C(t), lagging behind.  prg[w] = number of rotations column warp
            // w has applied (and, inside its own range, generated); every wait is on an earlier index, so no deadlock.
            // A warp issues in order, so every instruction in the driver loop costs time: addresses are running 32-bit
            // shared addresses and the operands of iteration t+1 are loaded during iteration t.
            volatile int* prg = ctl + 4;
            const int CW = c.nthr >> 6;                // column warps; the same number of row warps
            if (c.tid < 4) prg[c.tid] = l;
            __syncthreads();                           // (also: everyone has read H before the sweep writes)
            const int warp = c.tid >> 5, lane = c.tid & 31;
            const unsigned sH = kh_saddr(Hb), sQ = kh_saddr(Q);
            if (warp < CW) {
                const int w = warp, j = l + 32 * w + lane;
                if (l + 32 * w <= iact) {
                    const int jl = j < n ? j : n - 1;                   // clamped for the unconditional loads
                    const bool mine = (j <= iact);
                    const unsigned sP = kh_saddr(ctl + 4 + w);
                    kh_givens G = make_givens(f_first, g_first);       // G(l), computed redundantly by every column warp
                    unsigned a_t = sH + 16u * (unsigned)(ROWOFF(l) + jl);   // &H[t][j]
                    int rs = ROWSTEP(l);                                // elements from row t to row t+1
                    cd carry = kh_lds_cd(a_t);                          // H[t][j] before R(t)
                    cd h1 = kh_lds_cd(a_t + 16u * rs);                  // H[t+1][j]   (untouched by this sweep)
                    {   // R(l) on the own column (column l included: the shift makes its bottom entry the first bulge)
                        const cd top = G.c * carry + G.s * h1, bot = G.c * h1 - cconj(G.s) * carry;
                        if (mine) { kh_sts_cd(a_t, top); carry = bot; }
                    }
                    cd sub = mk(0, 0);
                    if (w == 0) { sub = kh_shfl_cd(carry, 0); if (lane == 0) { Q[l].c = G.c; Q[l].s = G.s; } }
                    __syncwarp();
                    if (lane == 0) kh_st_release_saddr(sP, l + 1);
                    const int tlast = (iact - 1 < l + 32 * w + 31) ? iact - 1 : l + 32 * w + 31;    // last rotation handled here
                    const int tgen = (w == 0) ? l + 1 : l + 32 * w;     // first rotation generated here
                    int t = l + 1;
                    a_t += 16u * rs; rs -= 1;
                    int spins = 0;
                    // ---- follower mode: rotations generated by the warps to the left
                    while (t < tgen && t <= tlast) {
                        const int wg = (t - l) >> 5;
                        kh_stress_pause(a.stress, t + 7 * w + b);
                        const int avail = kh_ld_acquire_saddr(kh_saddr(ctl + 4 + wg));
                        if (avail <= t) { if (++spins > a.spin_cap) { ctl[2] = 1; break; } continue; }
                        int tend = l + 32 * (wg + 1);                   // generation range of warp wg ends here
                        if (tend > avail) tend = avail;
                        if (tend > tgen) tend = tgen;
                        if (tend > tlast + 1) tend = tlast + 1;
                        h1 = kh_lds_cd(a_t + 16u * rs);
                        unsigned aq = sQ + (unsigned)sizeof(kh_qrot) * t;
                        for (; t < tend; ++t) {
                            const double cc = kh_lds_d(aq); const cd ss = kh_lds_cd(aq + 16);
                            const unsigned a_t1 = a_t + 16u * rs;
                            const cd h1n = kh_lds_cd(a_t1 + 16u * (rs - 1));   // next row, in flight while this rotation is applied
                            const cd top = cc * carry + ss * h1;
                            carry = cc * h1 - cconj(ss) * carry;
                            if (mine) kh_sts_cd(a_t, top);
                            h1 = h1n; a_t = a_t1; rs -= 1; aq += (unsigned)sizeof(kh_qrot);
                        }
                        G.c = kh_lds_d(aq - 80); G.s = kh_lds_cd(aq - 64); sub = kh_lds_cd(aq - 16);
                        __syncwarp();
                        if (lane == 0) kh_st_release_saddr(sP, t);
                    }
                    // ---- driver mode: this warp owns the corner column
                    if (t >= tgen && t <= tlast && ctl[2] == 0) {
                        unsigned a_hd = sH + 16u * (unsigned)(ROWOFF(t + 1) + t);          // &H[t+1][t]
                        unsigned aq = sQ + (unsigned)sizeof(kh_qrot) * t;
                        int lc = (t - l) & 31;                          // lane of the corner column t
                        cd hd = kh_lds_cd(a_hd);
                        h1 = kh_lds_cd(a_t + 16u * rs);
#pragma unroll 2
                        for (; t <= tlast; ++t) {
                            const unsigned a_t1 = a_t + 16u * rs;
                            a_hd += 16u * rs;
                            const cd hdn = kh_lds_cd(a_hd);             // operands of the next iteration, loaded ahead
                            const cd h1n = kh_lds_cd(a_t1 + 16u * (rs - 1));   // (row t+2 is at most one past the last row: padded)
                            const cd cb = kh_shfl_cd(carry, lc);                                      // H[t][t] after R(t-1)
                            const cd a1 = G.c * sub + cconj(G.s) * cb, b1 = G.c * cb - G.s * sub;      // C(t-1) on row t
                            const cd c1 = cconj(G.s) * hd, d1 = G.c * hd;                              // C(t-1) on row t+1: bulge
                            const double f2 = cabs2(a1), g2 = cabs2(c1);
                            kh_givens Gn;                                                              // G(t) annihilates it
                            if (f2 == 0.0 || g2 == 0.0) Gn = make_givens(a1, c1);                      // (warp-uniform, rare)
                            else Gn = make_givens_fast(a1, c1, f2, g2);
                            const cd ncs = cconj(Gn.s);
                            const cd top = Gn.c * carry + Gn.s * h1, bot = Gn.c * h1 - ncs * carry;    // R(t) on the own column
                            const cd newdiag = Gn.c * b1 + Gn.s * d1, nextsub = Gn.c * d1 - ncs * b1;
                            if (j > t && mine) { kh_sts_cd(a_t, top); carry = bot; }
                            if (lane == 0) {
                                kh_sts_d(aq, Gn.c); kh_sts_cd(aq + 16, Gn.s); kh_sts_cd(aq + 32, Gn.r);
                                kh_sts_cd(aq + 48, newdiag); kh_sts_cd(aq + 64, nextsub);
                            }
                            __syncwarp();
                            if (lane == 0) kh_st_release_saddr(sP, t + 1);
                            kh_stress_pause(a.stress, 3 * t + b);
                            G = Gn; sub = nextsub; hd = hdn; h1 = h1n; a_t = a_t1; rs -= 1;
                            aq += (unsigned)sizeof(kh_qrot); lc = (lc + 1) & 31;
                        }
                    }
                    // ---- the warp that owns column iact finishes row iact: C(iact-1) on its two entries
                    if (((iact - l) >> 5) == w && ctl[2] == 0) {
                        const cd cb = kh_shfl_cd(carry, (iact - l) & 31);
                        if (lane == 0) {
                            const int oi = ROWOFF(iact);
                            Hb[oi + iact - 1] = G.c * sub + cconj(G.s) * cb;
                            Hb[oi + iact] = G.c * cb - G.s * sub;
                        }
                    }
                }
            } else {
                // ---- row warps: C(t) on row r for t >= r (rows t+1, t+2 are handled inside the chain)
                const int r = l + (c.tid - 32 * CW);
                const int rmin = l + (warp - CW) * 32;
                const unsigned a_row = sH + 16u * (unsigned)((r < n) ? ROWOFF(r) : 0);
                cd car = mk(0, 0);
                if (rmin < iact) {
                    int t = rmin, spins = 0;
                    while (t < iact) {
                        const int wq = (t + 1 - l) >> 5;               // owner of column t+1 must have applied R(<= t) to it
                        kh_stress_pause(a.stress, t + 13 * warp + b);
                        const int avail = kh_ld_acquire_saddr(kh_saddr(ctl + 4 + wq));
                        if (avail <= t) { if (++spins > a.spin_cap) { ctl[2] = 1; break; } continue; }
                        int tend = l + 32 * (wq + 1) - 1;
                        if (tend > avail) tend = avail;
                        if (tend > iact) tend = iact;
                        unsigned aq = sQ + (unsigned)sizeof(kh_qrot) * t;
                        unsigned a_e = a_row + 16u * t;                  // &H[r][t]
                        for (; t < tend; ++t) {
                            const double cc = kh_lds_d(aq); const cd ss = kh_lds_cd(aq + 16);
                            if (r <= t) {
                                cd h0 = car;
                                if (r == t) { if (t > l) { h0 = kh_lds_cd(aq + 48); kh_sts_cd(a_e - 16, kh_lds_cd(aq + 32)); } else h0 = kh_lds_cd(a_e); }
                                const cd h1 = kh_lds_cd(a_e + 16);
                                kh_sts_cd(a_e, cc * h0 + cconj(ss) * h1);
                                car = cc * h1 - ss * h0;
                            }
                            aq += (unsigned)sizeof(kh_qrot); a_e += 16;
                        }
                    }
                    if (r < iact) Hb[ROWOFF(r) + iact] = car;
                }
            }
            __syncthreads();
            if (ctl[2] != 0) { fail = nfull + 1; break; }      // a consumer gave up waiting (spin_cap): the sweep is incomplete, stop here (uniform: read after the barrier)
        } else
#endif
        if (!PACKED) {
            c.sync();                               // everyone has read H before the sweep writes
            zqr_tiled_sweep(c, Hg, ldg, l, iact, f_first, g_first, Q, Hp);
        } else {
        kh_givens G = make_givens(f_first, g_first);
        c.sync();                                   // everyone has read H before the sweep writes
        for (int j = l + c.tid; j <= iact; j += c.nthr) {          // R(l)
            cd h0 = HQ(l, j), h1 = HQ(l + 1, j);
            HQ(l, j) = G.c * h0 + G.s * h1;
            HQ(l + 1, j) = G.c * h1 - cconj(G.s) * h0;
        }
        if (c.tid == 0) { Q[l].c = G.c; Q[l].s = G.s; }
        c.sync();
        // register-carried corner state: sub = H[k+1][k] after R(k); pend_* = entries of row k (its
        // sub-diagonal and diagonal after R(k)) that are known to every thread but not stored yet
        cd sub = HQ(l + 1, l);
        cd pend_sub = mk(0, 0), pend_diag = mk(0, 0);
        int o1 = ROWOFF(l + 1);                     // offset of row k+1 (running)
        const int m0 = l + c.tid, moff = ROWOFF(m0 < n ? m0 : 0);
        const bool one_idx = PACKED && c.nthr >= n;
        for (int k = l; k < iact; ++k) {
            const bool more = (k + 1 < iact);
            const int o2 = o1 + ROWSTEP(k + 1);     // offset of row k+2
            const cd cb = Hb[o1 + k + 1];
            const cd a1 = G.c * sub + cconj(G.s) * cb, b1 = G.c * cb - G.s * sub;     // C(k) on row k+1
            kh_givens Gn; Gn.c = 1.0; Gn.s = mk(0, 0); Gn.r = a1;
            cd newdiag = b1, nextsub = mk(0, 0);
            if (more) {
                const cd hd = Hb[o2 + k + 1];               // H[k+2][k] is zero: C(k) creates the bulge there
                const cd c1 = cconj(G.s) * hd, d1 = G.c * hd;
                Gn = make_givens(a1, c1);                   // G(k+1) annihilates the bulge
                newdiag = Gn.c * b1 + Gn.s * d1;            // H[k+1][k+1] after R(k+1)
                nextsub = Gn.c * d1 - cconj(Gn.s) * b1;     // H[k+2][k+1] after R(k+1)
            }
            // M(k) = C(k) on rows l..k  ||  R(k+1) on columns k+2..iact   (disjoint entries)
            for (int m = m0; m <= iact; m += c.nthr) {
                if (m <= k) {
                    cd* hr = Hb + (one_idx ? moff : ROWOFF(m)) + k;
                    cd h0 = hr[0], h1 = hr[1];
                    if (m == k && k > l) { h0 = pend_diag; hr[-1] = pend_sub; }   // row k: stored now, nobody reads it earlier
                    hr[0] = G.c * h0 + cconj(G.s) * h1;
                    hr[1] = G.c * h1 - G.s * h0;
                } else if (more && m < iact) {
                    const int j = m + 1;
                    cd h0 = Hb[o1 + j], h1 = Hb[o2 + j];
                    Hb[o1 + j] = Gn.c * h0 + Gn.s * h1;
                    Hb[o2 + j] = Gn.c * h1 - cconj(Gn.s) * h0;
                }
                if (one_idx) break;                  // packed path on the GPU: nthr >= n, one index per thread
            }
            if (more && c.tid == 0) { Q[k + 1].c = Gn.c; Q[k + 1].s = Gn.s; }
            pend_sub = Gn.r; pend_diag = newdiag; sub = nextsub; G = Gn; o1 = o2;
            c.sync();
        }
        if (c.tid == 0) { HQ(iact, iact - 1) = pend_sub; HQ(iact, iact) = pend_diag; }
        }
        QT_ADD(qt_sweep);
        // ---- delayed application of the sweep's rotations outside the window and to Z (no barriers inside)
        {
            const int nAbove = PACKED ? l : 0, nRight = n - 1 - iact;     // (the tiled sweep has already updated the rows above)
            int nZ = n, nR = nRight;
            if (rl) {
                // Z and the columns of H right of the window only ever RECEIVE rotations (no later sweep reads them), so
                // they are updated after the iteration by zrot_apply: only record the sweep.
                const int cnt = iact - l;
                if (nsw >= a.sw_cap || nrot + cnt > a.rot_cap) { fail = nfull + 2; break; }     // log full: reported like a convergence failure
                for (int k = l + c.tid; k < iact; k += c.nthr) {
                    double* rr = rl_rot + 3LL * (nrot + k - l);
                    rr[0] = Q[k].c; rr[1] = Q[k].s.x; rr[2] = Q[k].s.y;
                }
                if (c.tid == 0) { rl_sw[2 * nsw] = l | (iact << 16); rl_sw[2 * nsw + 1] = nrot; }
                nsw += 1; nrot += cnt;
                nZ = 0; nR = 0;
            }
            for (int t2 = c.tid; t2 < nAbove + nR + nZ; t2 += c.nthr) {
                if (t2 < nAbove) {                                  // rows above the window: column rotations
                    const int r = t2;
                    cd h0 = HQ(r, l);
                    for (int k = l; k < iact; ++k) {
                        cd h1 = HQ(r, k + 1);
                        HQ(r, k) = Q[k].c * h0 + cconj(Q[k].s) * h1;
                        h0 = Q[k].c * h1 - Q[k].s * h0;
                    }
                    HQ(r, iact) = h0;
                } else if (t2 < nAbove + nR) {                      // columns right of the window: row rotations
                    const int j = iact + 1 + (t2 - nAbove);
                    cd h0 = HQ(l, j);
                    for (int k = l; k < iact; ++k) {
                        cd h1 = HQ(k + 1, j);
                        HQ(k, j) = Q[k].c * h0 + Q[k].s * h1;
                        h0 = Q[k].c * h1 - cconj(Q[k].s) * h0;
                    }
                    HQ(iact, j) = h0;
                } else {                                            // Z columns l..iact (rows of Zt), coalesced over i
                    const int i = t2 - nAbove - nR;
                    cd z0 = ZT(l, i);
                    int k = l;
                    if (k + 8 <= iact) {                             // software pipeline: the next 8 rows of Zt are in flight
                        cd zn[8], zp[8];                             // while the current 8 rotations are applied (hides the L2 latency)
#pragma unroll
                        for (int u = 0; u < 8; ++u) zn[u] = ZT(k + 1 + u, i);
                        for (; k + 8 <= iact; k += 8) {
                            const bool nxt = (k + 16 <= iact);
                            if (nxt) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) zp[u] = ZT(k + 9 + u, i);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const double cc = Q[k + u].c; const cd ss = Q[k + u].s;
                                ZT(k + u, i) = cc * z0 + cconj(ss) * zn[u];
                                z0 = cc * zn[u] - ss * z0;
                            }
                            if (nxt) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) zn[u] = zp[u];
                            }
                        }
                    }
                    for (; k < iact; ++k) {
                        cd z1 = ZT(k + 1, i);
                        ZT(k, i) = Q[k].c * z0 + cconj(Q[k].s) * z1;
                        z0 = Q[k].c * z1 - Q[k].s * z0;
                    }
                    ZT(iact, i) = z0;
                }
            }
        }
        c.sync();
        QT_ADD(qt_delay);
#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
        qt_n += 1; qt_rot += iact - l;
#endif
    }
    c.sync();
    if (packed)
        for (int e = c.tid; e < n * n; e += c.nthr) {
            int i = e / n, j = e - i * n;
            Hg[(long long)i * ldg + j] = (j >= i - 1) ? HQ(i, j) : mk(0.0, 0.0);      // deflated sub-diagonals are exact zeros
        }
#if defined(KH_QR_TIMING) && !defined(KH_HOST_EMU)
    if (c.tid == 0 && b == 0) {
        kh_qr_dbg[0] += clock64() - qt0; kh_qr_dbg[1] += qt_scan; kh_qr_dbg[2] += qt_shift; kh_qr_dbg[3] += qt_sweep;
        kh_qr_dbg[4] += qt_delay; kh_qr_dbg[5] += qt_n; kh_qr_dbg[6] += qt_rot;
    }
#endif
    if (rl && c.tid == 0) { ((int*)rl)[0] = nsw; ((int*)rl)[1] = nrot; }
    if (fail == 0 && ctl[2] != 0) fail = nfull + 1;
    if (a.istate && c.tid == 0) a.istate[b] = fail ? -1 : iact;
    if (a.info && c.tid == 0 && (fail || a.phase == 0)) a.info[b] = fail;
#undef HQ
#undef ROWOFF
#undef ROWSTEP
#undef ZT
}

KH_DEV void zqr_packed_body(const Cta& c, const zgeev_args& a) { zqr_body_t<true>(c, a); }
KH_DEV void zqr_global_body(const Cta& c, const zgeev_args& a) { zqr_body_t<false>(c, a); }

// ============================================================================ 2b. replay of the rotation log
// Two CTAs per matrix.  by = 0: Zt (transposed Schur vectors) is staged in shared memory, thread i owns column i of Zt
// (= row i of Z) and applies every logged rotation to it.  by = 1: the triangular factor T is staged, thread j owns
// column j and applies the row rotations of every sweep whose window ended above row j (iact < j).  One LDS + one STS per
// rotation, the running entry stays in a register; the log is streamed through shared memory in chunks of whole
// sweeps (asynchronous copies, next chunk in flight while the current one is applied), one barrier per chunk.
struct zrot_args { int n; MatRef Zt, T; const double* rlog; long long rlog_stride; int sw_cap, chunk, cw; };
// Matrices beyond shared memory are replayed in column STRIPS of cw columns (grid.y = 2 * strips): the strip stays
// resident in shared memory for the whole log, so Z and T are read and written exactly once.
KH_DEV void zrot_apply_body(const Cta& c, const zrot_args& a) {
    const int n = a.n, b = c.bx, mode = c.by & 1, CH = a.chunk, cw = a.cw;
    const int c0 = (c.by >> 1) * cw, cols = (cw < n - c0) ? cw : n - c0;
    const MatRef M = mode ? a.T : a.Zt;
    cd* Mg = mat_ptr(M, b) + c0;
    const int ldm = M.ld;
    const double* rl = a.rlog + (long long)b * a.rlog_stride;
    const int nsw = ((const int*)rl)[0];
    const int* swg = (const int*)(rl + 2);
    const double* rot = rl + 2 + a.sw_cap;
    // shared: [Ms n x cw cd][rb 2 x CH x 4 dbl][sw 2 x sw_cap int]
    cd* Ms = (cd*)KH_SMEM(c);
    double* rb = (double*)(KH_SMEM(c) + (size_t)n * cw * sizeof(cd));
    int* sws = (int*)(KH_SMEM(c) + (size_t)n * cw * sizeof(cd) + (size_t)8 * CH * sizeof(double));
    if (nsw == 0 || cols <= 0) return;                               // uniform
    for (int e = c.tid; e < 2 * nsw; e += c.nthr) sws[e] = swg[e];
    kh_stage_rows(c, Ms, cw, Mg, ldm, n, cols, (unsigned long long*)(sws + 2 * a.sw_cap));      // strip rows by TMA bulk copies
    const int nrot = sws[2 * nsw - 1] + ((sws[2 * nsw - 2] >> 16) - (sws[2 * nsw - 2] & 0xffff));
    // chunk [s0, s1): whole sweeps, at most CH rotations (every sweep has fewer than n <= CH rotations)
    auto chunk_end = [&](int s0) {
        int s1 = s0 + 1;
        while (s1 < nsw && ((s1 + 1 < nsw ? sws[2 * s1 + 3] : nrot) - sws[2 * s0 + 1]) <= CH) ++s1;
        return s1;
    };
    auto stage = [&](int s0, int s1, double* dst) {
        const int r0 = sws[2 * s0 + 1], r1 = (s1 < nsw) ? sws[2 * s1 + 1] : nrot;
        const double* src = rot + 3LL * r0;
        for (int e = c.tid; e < 3 * (r1 - r0); e += c.nthr) {
            const int q = e / 3, comp = e - 3 * q;
#ifdef KH_HOST_EMU
            dst[4 * q + comp] = src[e];
#else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst + 4 * q + comp)), "l"(src + e));
#endif
        }
    };
    int s0 = 0, s1 = chunk_end(0), buf = 0;
    stage(s0, s1, rb);
#ifndef KH_HOST_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    c.sync();
    while (s0 < nsw) {
        const int s2 = (s1 < nsw) ? chunk_end(s1) : s1;
        if (s1 < nsw) stage(s1, s2, rb + (buf ^ 1) * 4 * CH);
        const double* cur = rb + buf * 4 * CH;
        const int rbase = sws[2 * s0 + 1];
        for (int i = c.tid; i < cols; i += c.nthr) {
            for (int sw = s0; sw < s1; ++sw) {
                const int l = sws[2 * sw] & 0xffff, iact = sws[2 * sw] >> 16;
                if (mode && iact >= c0 + i) continue;
                const double* q = cur + 4 * (sws[2 * sw + 1] - rbase);
                cd* zp = Ms + (long long)l * cw + i;
                cd z0 = zp[0];
                int k = l;
                if (mode == 0) {
                    for (; k + 4 <= iact; k += 4, q += 16, zp += 4 * cw) {      // rows k+1..k+4 are read before this sweep writes them
                        const cd z1 = zp[cw], z2 = zp[2 * cw], z3 = zp[3 * cw], z4 = zp[4 * cw];
                        const double c0_ = q[0], c1 = q[4], c2 = q[8], c3 = q[12];
                        const cd t0 = mk(q[1], q[2]), t1 = mk(q[5], q[6]), t2 = mk(q[9], q[10]), t3 = mk(q[13], q[14]);
                        zp[0] = c0_ * z0 + cconj(t0) * z1; z0 = c0_ * z1 - t0 * z0;
                        zp[cw] = c1 * z0 + cconj(t1) * z2; z0 = c1 * z2 - t1 * z0;
                        zp[2 * cw] = c2 * z0 + cconj(t2) * z3; z0 = c2 * z3 - t2 * z0;
                        zp[3 * cw] = c3 * z0 + cconj(t3) * z4; z0 = c3 * z4 - t3 * z0;
                    }
                    for (; k < iact; ++k, q += 4, zp += cw) {
                        const cd z1 = zp[cw];
                        const double cc = q[0]; const cd ss = mk(q[1], q[2]);
                        zp[0] = cc * z0 + cconj(ss) * z1; z0 = cc * z1 - ss * z0;
                    }
                } else {
                    for (; k + 4 <= iact; k += 4, q += 16, zp += 4 * cw) {
                        const cd z1 = zp[cw], z2 = zp[2 * cw], z3 = zp[3 * cw], z4 = zp[4 * cw];
                        const double c0_ = q[0], c1 = q[4], c2 = q[8], c3 = q[12];
                        const cd t0 = mk(q[1], q[2]), t1 = mk(q[5], q[6]), t2 = mk(q[9], q[10]), t3 = mk(q[13], q[14]);
                        zp[0] = c0_ * z0 + t0 * z1; z0 = c0_ * z1 - cconj(t0) * z0;
                        zp[cw] = c1 * z0 + t1 * z2; z0 = c1 * z2 - cconj(t1) * z0;
                        zp[2 * cw] = c2 * z0 + t2 * z3; z0 = c2 * z3 - cconj(t2) * z0;
                        zp[3 * cw] = c3 * z0 + t3 * z4; z0 = c3 * z4 - cconj(t3) * z0;
                    }
                    for (; k < iact; ++k, q += 4, zp += cw) {
                        const cd z1 = zp[cw];
                        const double cc = q[0]; const cd ss = mk(q[1], q[2]);
                        zp[0] = cc * z0 + ss * z1; z0 = cc * z1 - cconj(ss) * z0;
                    }
                }
                zp[0] = z0;
            }
        }
#ifndef KH_HOST_EMU
        asm volatile("cp.async.wait_all;" ::: "memory");
#endif
        c.sync();
        s0 = s1; s1 = s2; buf ^= 1;
    }
    for (int e = c.tid; e < n * cols; e += c.nthr) { const int k = e / cols, i = e - k * cols; Mg[(long long)k * ldm + i] = Ms[k * cw + i]; }
}
// strip width: the whole matrix when it fits beside a chunk buffer of >= n rotations, else equal strips of <= 64 columns
static inline int zrot_strip(int n, int sw_cap) {
    const long long fixed = 2LL * sw_cap * (long long)sizeof(int) + 64, ch_min = 2LL * 4 * (long long)sizeof(double) * (n > 256 ? n : 256);
    if ((long long)n * n * (long long)sizeof(cd) + fixed + 2LL * 4 * (long long)sizeof(double) * n <= (long long)KH_SMEM_MAX) return n;
    long long room = (long long)KH_SMEM_MAX - fixed - ch_min;
    if (room <= 0) return 0;
    long long cw = room / ((long long)n * (long long)sizeof(cd));
    if (cw > 64) cw = 64;
    if (cw < 1) return 0;
    const int strips = (int)((n + cw - 1) / cw);
    return (n + strips - 1) / strips;
}
static inline int zrot_chunk(int n, int sw_cap, int cw) {             // rotations per staged chunk that fit beside the strip
    const long long room = (long long)KH_SMEM_MAX - (long long)n * cw * (long long)sizeof(cd) - 2LL * sw_cap * (long long)sizeof(int) - 64;
    long long ch = room / (2 * 4 * (long long)sizeof(double));
    return ch > 4096 ? 4096 : (int)ch;
}
static inline size_t zrot_smem_bytes(int n, int sw_cap, int chunk, int cw) { return (size_t)n * cw * sizeof(cd) + (size_t)8 * chunk * sizeof(double) + (size_t)2 * sw_cap * sizeof(int) + 16; }      // (+16: mbarrier of the staging copy)
// work space (doubles per matrix) the log needs for the given capacities
static inline long long zgeev_rlog_doubles(int rot_cap, int sw_cap) { return 2LL + sw_cap + 3LL * rot_cap; }

// ============================================================================ 3. eigenvectors of the triangular factor
KH_DEV void ztrevc_body(const Cta& c, const zgeev_args& a) {
    const int n = a.n, b = c.bx;
    const cd* T = mat_ptr(a.Hw, b);
    cd* X = mat_ptr(a.X, b);
    const int ldt = a.Hw.ld, ldx = a.X.ld;
    const double smlnum = ZGEEV_SAFMIN * ((double)n / ZGEEV_EPS);
#define TT(i, j) T[(long long)(i) * ldt + (j)]
#define XX(i, j) X[(long long)(i) * ldx + (j)]
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int i = e / n, k = e - i * n;
        XX(i, k) = (i < k) ? -TT(i, k) : mk(i == k ? 1.0 : 0.0, 0.0);
    }
    c.sync();
    for (int k = c.tid; k < n; k += c.nthr) {
        cd tkk = TT(k, k);
        double smin = fmax(ZGEEV_EPS * cabs1(tkk), smlnum);
        for (int j = n - 2; j >= 0; --j) {
            if (j < k) {
                cd d = TT(j, j) - tkk;
                if (cabs1(d) < smin) d = mk(smin, 0.0);
                cd xj = XX(j, k) / d;
                XX(j, k) = xj;
                for (int i = 0; i < j; ++i) { cd v = XX(i, k); cfms(v, xj, TT(i, j)); XX(i, k) = v; }
            }
        }
        double emax = 0.0;
        for (int i = 0; i <= k; ++i) emax = fmax(emax, cabs1(XX(i, k)));
        double r = 1.0 / emax;
        for (int i = 0; i <= k; ++i) XX(i, k) = r * XX(i, k);
    }
#undef TT
#undef XX
}

#ifndef KH_ZQT_SMEM_PAD
#define KH_ZQT_SMEM_PAD 0             /* extra dynamic shared memory: caps the CTAs per SM (L2 residency of the matrices in flight) */
#endif
static inline size_t zhess_smem_bytes(int n, int ld_s, int use_smem) {
    size_t s = (size_t)2 * n * sizeof(cd) + 192 * sizeof(double) + (size_t)4 * n * sizeof(double) + 32;
    if (use_smem) s += (size_t)n * ld_s * sizeof(cd);
    return s;
}
static inline size_t zqr_smem_bytes(int n, int use_smem) {
    size_t s = (size_t)n * sizeof(kh_qrot) + 16 * sizeof(int) + 16;
    if (use_smem) s += (size_t)hp_size(n) * sizeof(cd);
    else s += (size_t)(ZQT_W + 2) * (ZQT_LD + 2) * sizeof(cd) + KH_ZQT_SMEM_PAD;
    return s;
}

#ifndef KH_ZQT_THREADS
#define KH_ZQT_THREADS 128       /* tiled QR (global-memory Hessenberg matrix): threads per CTA, CTAs per SM */
#define KH_ZQT_MINB 4
#endif
#ifndef KH_QR_THREADS
#define KH_QR_THREADS(n) ((n) <= 64 ? 128 : 256)
#endif
static inline int zgeev_launch(kh_stream_t st, int batch, zgeev_args a) {
    if (batch <= 0 || a.n <= 0) return 0;
    const int n = a.n;
    const double work = 100.0 * n * n * n * batch;          // nominal zgeev count, SURVEY.md 8(d)
    a.ld_s = n | 1;
    a.use_smem = zhess_smem_bytes(n, a.ld_s, 1) <= (size_t)KH_SMEM_MAX;
    int e;
    const bool tiled_hess = !a.use_smem && a.rlog && a.rlog_stride / 2 >= zhb_work_cd(n);
    if (tiled_hess) e = zhess_blocked_launch(st, batch, n, a.A, a.Hw, a.Zt, a.X, (cd*)a.rlog, a.rlog_stride / 2, a.scale, a.scale_stride, a.tau, a.tau_stride, 0.25 * work);
    else if (a.use_smem) e = kh_launch<zgeev_args, zhessz_body>(dim3(batch), n <= 42 ? 128 : (n <= 85 ? 256 : 384), zhess_smem_bytes(n, a.ld_s, 1), st, a, "zgeev_hess", 0.25 * work);
    else return KH_ENOMEM_ZGEEV;            // beyond shared memory the blocked reduction needs the log work space (kh_zgeev_work_bytes provides it)
    if (e) return e;
    zgeev_args q = a;
    {   // test switches of the relay sweep's flag protocol (tests/test_parity.py::test_qr_flag_protocol_*)
        const char* e1 = getenv("KH_QR_SPIN_CAP"); const char* e2 = getenv("KH_QR_STRESS");
        if (e1 && atoi(e1) > 0) q.spin_cap = atoi(e1);
        if (e2) q.stress = atoi(e2);
    }
    q.use_smem = zqr_smem_bytes(n, 1) <= (size_t)KH_SMEM_MAX;
    const int zcw = (q.rlog && q.sw_cap >= 1) ? zrot_strip(n, q.sw_cap) : 0;
    if (!(q.rlog && q.rot_cap >= n && q.sw_cap >= 1 && q.sw_cap < 65536 && n < 65536 && zcw > 0 &&
          zrot_chunk(n, q.sw_cap, zcw) >= n)) q.rlog = nullptr;
    if (q.use_smem && q.rlog && q.istate) {
        // Phased iteration on the shrinking leading block: (na, launch shape) chosen so that 2 / 4 / 8 CTAs share an SM.
        auto fit = [&](size_t budget) { int m = n; while (m > 8 && zqr_smem_bytes(m, 1) > budget) --m; return m; };
        const int na2 = fit((size_t)KH_SMEM_MAX / 4 - 1024), na3 = fit((size_t)KH_SMEM_MAX / 8 - 1024);
        int phase = 0;
        auto run = [&](int na, int stop) {
            zgeev_args r = q;
            r.na = na; r.iact_stop = stop; r.phase = phase++;
            const size_t sm = zqr_smem_bytes(na, 1);
            const int thr = na <= 64 ? 128 : (na <= 96 ? 192 : 256);
            if (na <= na3 && na <= 64) return kh_launch<zgeev_args, zqr_packed_body, 128, 8>(dim3(batch), thr, sm, st, r, "zgeev_qr", phase == 1 ? 0.5 * work : 0.0);
            if (na <= na2 && na <= 96) return kh_launch<zgeev_args, zqr_packed_body, 192, 4>(dim3(batch), thr, sm, st, r, "zgeev_qr", phase == 1 ? 0.5 * work : 0.0);
            return kh_launch<zgeev_args, zqr_packed_body, 256, 2>(dim3(batch), thr, sm, st, r, "zgeev_qr", phase == 1 ? 0.5 * work : 0.0);
        };
        if (n > na2 + 8) { e = run(n, na2); if (e) return e; }
        if (n > na3 + 8) { e = run(n < na2 ? n : (phase ? na2 : n), na3); if (e) return e; }
        e = run(phase ? na3 : n, 0);
    }
    else if (q.use_smem) e = kh_launch<zgeev_args, zqr_packed_body, 256, 2>(dim3(batch), KH_QR_THREADS(n), zqr_smem_bytes(n, 1), st, q, "zgeev_qr", 0.5 * work);
    else e = kh_launch<zgeev_args, zqr_global_body, KH_ZQT_THREADS, KH_ZQT_MINB>(dim3(batch), KH_ZQT_THREADS, zqr_smem_bytes(n, 0), st, q, "zgeev_qr", 0.5 * work);
    if (e) return e;
    if (q.rlog) {
        const int ch = zrot_chunk(n, q.sw_cap, zcw), strips = (n + zcw - 1) / zcw;
        zrot_args z{n, a.Zt, a.Hw, q.rlog, q.rlog_stride, q.sw_cap, ch, zcw};
        e = kh_launch<zrot_args, zrot_apply_body>(dim3(batch, 2 * strips), zcw <= 32 ? 32 : (zcw <= 64 ? 64 : 128), zrot_smem_bytes(n, q.sw_cap, ch, zcw), st, z, "zgeev_zrot", 0.0);
        if (e) return e;
    }
    return kh_launch<zgeev_args, ztrevc_body>(dim3(batch), n <= 128 ? 128 : 256, 0, st, a, "zgeev_trevc", 0.25 * work);
}
