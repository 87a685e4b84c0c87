// RCWA-specific kernels around the dense primitives (all complex128, one batch element = one
// (wavelength, k-point) solve).  Notation follows SURVEY.md §8: N harmonics, n = 2N.
//
// "BD" = a 2N x 2N matrix made of 2x2 blocks of N-diagonals (what uniform layers and half spaces
// produce).  It is stored compactly as four N-vectors (a, b; c, d) = one 2x2 complex matrix per
// harmonic, so BD (*) BD star products are O(N) instead of O(n^3).
#pragma once
#include "kh_common.cuh"

#define KH_TWO_PI 6.283185307179586476925286766559

struct m22 { cd a, b, c, d; };
KH_HD m22 m22_mul(const m22& p, const m22& q) {
    m22 r;
    r.a = p.a * q.a + p.b * q.c; r.b = p.a * q.b + p.b * q.d;
    r.c = p.c * q.a + p.d * q.c; r.d = p.c * q.b + p.d * q.d;
    return r;
}
KH_HD m22 m22_add(const m22& p, const m22& q) { m22 r; r.a = p.a + q.a; r.b = p.b + q.b; r.c = p.c + q.c; r.d = p.d + q.d; return r; }
KH_HD m22 m22_sub(const m22& p, const m22& q) { m22 r; r.a = p.a - q.a; r.b = p.b - q.b; r.c = p.c - q.c; r.d = p.d - q.d; return r; }
KH_HD m22 m22_scale(cd s, const m22& p) { m22 r; r.a = s * p.a; r.b = s * p.b; r.c = s * p.c; r.d = s * p.d; return r; }
KH_HD m22 m22_eye() { m22 r; r.a = mk(1, 0); r.b = mk(0, 0); r.c = mk(0, 0); r.d = mk(1, 0); return r; }
KH_HD m22 m22_inv(const m22& p) {
    // Gaussian elimination with row pivoting (what LAPACK does on each decoupled 2x2 block)
    m22 r;
    if (cabs1(p.a) >= cabs1(p.c)) {
        cd l = p.c / p.a;                 // eliminate c
        cd u22 = p.d - l * p.b;
        cd i22 = crecip(u22), i11 = crecip(p.a);
        // inverse of [[a, b], [0, u22]] times [[1, 0], [-l, 1]]
        cd t = -(p.b * i22) * i11;        // upper-right of U^-1
        r.a = i11 - t * l; r.b = t;
        r.c = -(i22 * l); r.d = i22;
    } else {
        cd l = p.a / p.c;                 // rows swapped: [[c, d], [a, b]]
        cd u22 = p.b - l * p.d;
        cd i22 = crecip(u22), i11 = crecip(p.c);
        cd t = -(p.d * i22) * i11;
        // (P A)^-1 = [[i11 - t*l, t], [-i22*l, i22]] ; A^-1 = (P A)^-1 P  -> swap columns
        r.b = i11 - t * l; r.a = t;
        r.d = -(i22 * l); r.c = i22;
    }
    return r;
}

// ------------------------------------------------------------------ k-vectors (expansion.py:43-50)
struct kvec_args {
    int B, N;
    const double* wl;        // [B]
    const cd* kp;            // [B][2]
    const double* g;         // [2][N]
    cd* Kx; cd* Ky;          // [B][N]
    double* k0;              // [B]
};
KH_DEV void kvec_body(const Cta& c, const kvec_args& a) {
    const int b = c.bx;
    const double k0 = KH_TWO_PI / a.wl[b];
    if (c.tid == 0) a.k0[b] = k0;
    const cd kpx = a.kp[2 * b], kpy = a.kp[2 * b + 1];
    for (int g = c.tid; g < a.N; g += c.nthr) {
        a.Kx[(long long)b * a.N + g] = mk((kpx.x + a.g[g]) / k0, kpx.y / k0);
        a.Ky[(long long)b * a.N + g] = mk((kpy.x + a.g[a.N + g]) / k0, kpy.y / k0);
    }
}

// free-space branch of kz (alternative.py:92-96, 146-149): sign of Re(kz^2) picks the root
KH_HD cd kz_branch(cd arg) {
    if (arg.x < 0.0) { cd s = csqrt_(-arg); return mk(s.y, -s.x); }     // -i * sqrt(-arg)
    return csqrt_(arg);
}
KH_HD cd times_i(cd z) { return mk(-z.y, z.x); }

// V-type 2x2 block  Q(eps)/lambda  for one harmonic:  [[KxKy, eps-Kx^2],[Ky^2-eps, -KyKx]] / lam
KH_HD m22 q_over_lam(cd kx, cd ky, cd eps, cd lam) {
    m22 q;
    q.a = (kx * ky) / lam; q.b = (eps - kx * kx) / lam;
    q.c = (ky * ky - eps) / lam; q.d = (-(ky * kx)) / lam;
    return q;
}
KH_HD m22 v0_block(cd kx, cd ky) {                                      // alternative.py:84-99
    cd one = mk(1, 0);
    cd lam0 = times_i(kz_branch(one - kx * kx - ky * ky));
    return q_over_lam(kx, ky, one, lam0);
}

// ------------------------------------------------------------------ analytic BD layers
// kinds follow khepri/layer.py:19-24 (Formulation): 0 uniform, 3 half-space incidence, 4 half-space emergence
struct bd_layer_args {
    int B, N;
    int kind;
    cd eps; double depth;
    const cd* Kx; const cd* Ky; const double* k0;
    cd* S;            // [B][4 blocks][4 entries a,b,c,d][N]
    cd* V;            // optional [B][4][N]: the layer's V block (fields); null otherwise
    cd* lam;          // optional [B][N]
};
KH_DEV void bd_store(cd* S, long long base, int N, int blk, int g, const m22& m) {
    cd* p = S + base + (long long)blk * 4 * N + g;
    p[0] = m.a; p[N] = m.b; p[2 * N] = m.c; p[3 * N] = m.d;
}
KH_DEV m22 bd_load(const cd* S, long long base, int N, int blk, int g) {
    const cd* p = S + base + (long long)blk * 4 * N + g;
    m22 m; m.a = p[0]; m.b = p[N]; m.c = p[2 * N]; m.d = p[3 * N];
    return m;
}
KH_DEV void bd_layer_body(const Cta& c, const bd_layer_args& a) {
    const int b = c.bx, N = a.N;
    const long long base = (long long)b * 16 * N;
    const cd one = mk(1, 0);
    for (int g = c.tid; g < N; g += c.nthr) {
        cd kx = a.Kx[(long long)b * N + g], ky = a.Ky[(long long)b * N + g];
        m22 V0 = v0_block(kx, ky);
        m22 S11, S12, S21, S22, V;
        cd lam;
        if (a.kind == 0) {
            // solve_uniform_layer + build_scatmat (alternative.py:130-156, 181-195), W = I
            lam = times_i(kz_branch(a.eps - kx * kx - ky * ky));
            m22 q;
            cd ie = crecip(a.eps);
            q.a = a.eps * (ie * (kx * ky)); q.b = a.eps * (ie * (a.eps - kx * kx));
            q.c = a.eps * (ie * (ky * ky - a.eps)); q.d = a.eps * (ie * (-(ky * kx)));
            V.a = q.a / lam; V.b = q.b / lam; V.c = q.c / lam; V.d = q.d / lam;
            m22 t2 = m22_mul(m22_inv(V), V0);
            m22 A = m22_add(m22_eye(), t2), Bm = m22_sub(m22_eye(), t2);
            cd X = cexp_((-a.depth * a.k0[b]) * lam);
            m22 Ai = m22_inv(A);
            m22 XB = m22_scale(X, Bm);
            m22 XBAiX = m22_scale(X, m22_mul(XB, Ai));            // X B A^-1 X
            m22 T = m22_sub(A, m22_mul(XBAiX, Bm));
            m22 Ti = m22_inv(T);
            S11 = m22_mul(Ti, m22_sub(m22_mul(XBAiX, A), Bm));
            S12 = m22_mul(Ti, m22_scale(X, m22_sub(A, m22_mul(Bm, m22_mul(Ai, Bm)))));
            S21 = S12; S22 = S11;
        } else {
            // scattering_reflection / scattering_transmission (alternative.py:32-82), W = I
            cd kz = cconj(csqrt_(a.eps - kx * kx - ky * ky));
            lam = times_i(kz);
            V = q_over_lam(kx, ky, a.eps, lam);
            m22 t2 = m22_mul(m22_inv(V0), V);
            m22 A = m22_add(m22_eye(), t2), Bm = m22_sub(m22_eye(), t2);
            m22 Ai = m22_inv(A);
            m22 AiB = m22_mul(Ai, Bm);
            m22 mAiB = m22_scale(mk(-1, 0), AiB);
            m22 twoAi = m22_scale(mk(2, 0), Ai);
            m22 half = m22_scale(mk(0.5, 0), m22_sub(A, m22_mul(Bm, AiB)));
            m22 BAi = m22_mul(Bm, Ai);
            if (a.kind == 3) { S11 = mAiB; S12 = twoAi; S21 = half; S22 = BAi; }
            else { S11 = BAi; S12 = half; S21 = twoAi; S22 = mAiB; }
        }
        (void)one;
        bd_store(a.S, base, N, 0, g, S11); bd_store(a.S, base, N, 1, g, S12);
        bd_store(a.S, base, N, 2, g, S21); bd_store(a.S, base, N, 3, g, S22);
        if (a.V) { cd* p = a.V + (long long)b * 4 * N + g; p[0] = V.a; p[N] = V.b; p[2 * N] = V.c; p[3 * N] = V.d; }
        if (a.lam) a.lam[(long long)b * N + g] = lam;
    }
}

// identity S-matrix in BD form (alternative.py:220-232)
struct bd_identity_args { int B, N; cd* S; };
KH_DEV void bd_identity_body(const Cta& c, const bd_identity_args& a) {
    const long long base = (long long)c.bx * 16 * a.N;
    m22 z; z.a = z.b = z.c = z.d = mk(0, 0);
    for (int g = c.tid; g < a.N; g += c.nthr) {
        bd_store(a.S, base, a.N, 0, g, z); bd_store(a.S, base, a.N, 1, g, m22_eye());
        bd_store(a.S, base, a.N, 2, g, m22_eye()); bd_store(a.S, base, a.N, 3, g, z);
    }
}

// BD (*) BD Redheffer star product (alternative.py:19-30), per harmonic
struct bd_star_args { int B, N; const cd* SA; const cd* SB; cd* SO; };
KH_DEV void bd_star_body(const Cta& c, const bd_star_args& a) {
    const int N = a.N;
    const long long base = (long long)c.bx * 16 * N;
    for (int g = c.tid; g < N; g += c.nthr) {
        m22 A11 = bd_load(a.SA, base, N, 0, g), A12 = bd_load(a.SA, base, N, 1, g);
        m22 A21 = bd_load(a.SA, base, N, 2, g), A22 = bd_load(a.SA, base, N, 3, g);
        m22 B11 = bd_load(a.SB, base, N, 0, g), B12 = bd_load(a.SB, base, N, 1, g);
        m22 B21 = bd_load(a.SB, base, N, 2, g), B22 = bd_load(a.SB, base, N, 3, g);
        m22 Di = m22_inv(m22_sub(m22_eye(), m22_mul(B11, A22)));
        m22 Fi = m22_inv(m22_sub(m22_eye(), m22_mul(A22, B11)));
        m22 S11 = m22_add(A11, m22_mul(m22_mul(A12, m22_mul(Di, B11)), A21));
        m22 S12 = m22_mul(A12, m22_mul(Di, B12));
        m22 S21 = m22_mul(B21, m22_mul(Fi, A21));
        m22 S22 = m22_add(B22, m22_mul(m22_mul(B21, m22_mul(Fi, A22)), B12));
        bd_store(a.SO, base, N, 0, g, S11); bd_store(a.SO, base, N, 1, g, S12);
        bd_store(a.SO, base, N, 2, g, S21); bd_store(a.SO, base, N, 3, g, S22);
    }
}

// BD -> dense [B][4][n][n]
struct bd_expand_args { int B, N; const cd* S; cd* D; };
KH_DEV void bd_expand_body(const Cta& c, const bd_expand_args& a) {
    const int N = a.N, n = 2 * N, blk = c.by;
    const cd* s = a.S + (long long)c.bx * 16 * N + (long long)blk * 4 * N;
    cd* d = a.D + ((long long)c.bx * 4 + blk) * n * n;
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int i = e / n, j = e - i * n;
        int gi = i < N ? i : i - N, gj = j < N ? j : j - N;
        cd v = mk(0, 0);
        if (gi == gj) v = s[((i >= N) * 2 + (j >= N)) * N + gi];
        d[e] = v;
    }
}

// dense (x) BD products, O(n^2):  out = alpha * (BD . M | M . BD) + beta * Cin + diag * I + expand(addbd)
// (used by the star products whose left or right operand is a uniform layer / half space)
struct bdmul_args {
    int B, N, side;                 // side 0: BD . M ; side 1: M . BD
    const cd* bd; int blk;          // BD table [B][4][4][N], block index
    MatRef M, Cin, out;             // Cin.p may be null
    const cd* addbd; int addblk;    // optional BD block added to the result
    double alpha, beta, diag;
    int csel, c0, c1;               // csel: only the output columns c0, c1 are formed (flux columns of the star chain), O(n) work
};
KH_DEV void bdmul_body(const Cta& c, const bdmul_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const cd* bd = a.bd + (long long)b * 16 * N + (long long)a.blk * 4 * N;
    const cd* ad = a.addbd ? a.addbd + (long long)b * 16 * N + (long long)a.addblk * 4 * N : (const cd*)0;
    const cd* M = mat_ptr(a.M, b);
    const cd* Cin = mat_ptr(a.Cin, b);
    cd* out = mat_ptr(a.out, b);
    const int rows_per = a.csel ? n : (n + 3) / 4, r0 = c.by * rows_per, r1 = (r0 + rows_per < n) ? r0 + rows_per : n;      // csel: one CTA per solve
    const int width = a.csel ? 2 : n;
    for (int e = r0 * width + c.tid; e < r1 * width; e += c.nthr) {
        const int i = e / width, jj = e - i * width, j = a.csel ? (jj ? a.c1 : a.c0) : jj;
        const int hi = i >= N, gi = i - hi * N, hj = j >= N, gj = j - hj * N;
        cd v;
        if (a.side == 0) v = bd[(hi * 2 + 0) * N + gi] * M[(long long)gi * a.M.ld + j] + bd[(hi * 2 + 1) * N + gi] * M[(long long)(N + gi) * a.M.ld + j];
        else v = M[(long long)i * a.M.ld + gj] * bd[(0 * 2 + hj) * N + gj] + M[(long long)i * a.M.ld + N + gj] * bd[(1 * 2 + hj) * N + gj];
        v = a.alpha * v;
        if (Cin) v = v + a.beta * Cin[(long long)i * a.Cin.ld + j];
        if (i == j) v.x += a.diag;
        if (ad && gi == gj) v = v + ad[(hi * 2 + hj) * N + gi];
        out[(long long)i * a.out.ld + j] = v;
    }
}

// ------------------------------------------------------------------ patterned layer: P, Q (alternative.py:158-171)
struct pq_args {
    int B, N;
    const cd* C; const cd* IC;          // [N][N], shared by the batch
    const cd* Kx; const cd* Ky;         // [B][N]
    cd* P; cd* Q;                       // [B][n][n]
};
KH_DEV void pq_body(const Cta& c, const pq_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const cd* kx = a.Kx + (long long)b * N;
    const cd* ky = a.Ky + (long long)b * N;
    cd* P = a.P + (long long)b * n * n;
    cd* Q = a.Q + (long long)b * n * n;
    for (int e = c.tid; e < N * N; e += c.nthr) {
        int r = e / N, q = e - r * N;
        cd ic = a.IC[e], cc = a.C[e];
        double dl = (r == q) ? 1.0 : 0.0;
        cd icky = ic * ky[q], ickx = ic * kx[q];
        cd p11 = kx[r] * icky, p12 = mk(dl, 0) - kx[r] * ickx;
        cd p21 = ky[r] * icky - mk(dl, 0), p22 = -(ky[r] * ickx);
        P[(long long)r * n + q] = p11; P[(long long)r * n + N + q] = p12;
        P[(long long)(N + r) * n + q] = p21; P[(long long)(N + r) * n + N + q] = p22;
        cd z = mk(0, 0);
        cd q11 = z, q12 = cc, q21 = -cc, q22 = z;
        if (r == q) {
            q11 = kx[r] * ky[r]; q12 = cc - kx[r] * kx[r];
            q21 = ky[r] * ky[r] - cc; q22 = -(ky[r] * kx[r]);
        }
        Q[(long long)r * n + q] = q11; Q[(long long)r * n + N + q] = q12;
        Q[(long long)(N + r) * n + q] = q21; Q[(long long)(N + r) * n + N + q] = q22;
    }
}

// lambda = sqrt(lambda^2 + 0j), X = exp(-lambda d k0)   (alternative.py:173, 186)
struct lam_args { int B, n; double depth; const cd* w; const double* k0; cd* lam; cd* xexp; cd* ilam; };
KH_DEV void lam_body(const Cta& c, const lam_args& a) {
    const int b = c.bx;
    for (int i = c.tid; i < a.n; i += c.nthr) {
        cd w = a.w[(long long)b * a.n + i];
        cd l = csqrt_(mk(w.x + 0.0, w.y + 0.0));
        a.lam[(long long)b * a.n + i] = l;
        a.xexp[(long long)b * a.n + i] = cexp_((-a.depth * a.k0[b]) * l);
        if (a.ilam) a.ilam[(long long)b * a.n + i] = crecip(l);
    }
}

// A = W^-1 + V^-1 V0 ; B = W^-1 - V^-1 V0 ; XB = X B ; XA = X A     (alternative.py:182-186; W0 = I)
struct ab_args {
    int B, N;
    const cd* Winv; const cd* Vinv;      // [B][n][n]
    const cd* Kx; const cd* Ky; const cd* xexp;
    cd* A; cd* Bm; cd* XB; cd* XA;       // [B][n][n] each
};
KH_DEV void ab_body(const Cta& c, const ab_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const long long off = (long long)b * n * n;
    const cd* kx = a.Kx + (long long)b * N;
    const cd* ky = a.Ky + (long long)b * N;
    const cd* x = a.xexp + (long long)b * n;
    for (int e = c.tid; e < n * N; e += c.nthr) {
        int i = e / N, g = e - i * N;
        m22 V0 = v0_block(kx[g], ky[g]);
        cd v1 = a.Vinv[off + (long long)i * n + g], v2 = a.Vinv[off + (long long)i * n + N + g];
        cd t1 = v1 * V0.a + v2 * V0.c, t2 = v1 * V0.b + v2 * V0.d;
        cd w1 = a.Winv[off + (long long)i * n + g], w2 = a.Winv[off + (long long)i * n + N + g];
        cd a1 = w1 + t1, a2 = w2 + t2, b1 = w1 - t1, b2 = w2 - t2;
        long long o1 = off + (long long)i * n + g, o2 = o1 + N;
        a.A[o1] = a1; a.A[o2] = a2; a.Bm[o1] = b1; a.Bm[o2] = b2;
        a.XB[o1] = x[i] * b1; a.XB[o2] = x[i] * b2;
        a.XA[o1] = x[i] * a1; a.XA[o2] = x[i] * a2;
    }
}

// PV0 = P . V0 (V0 = free-space H modes, 2x2 blocks of diagonals).  Since Omega^2 = P Q = W L^2 W^-1 and
// V = Q W L^-1, one has V^-1 = L^-1 W^-1 P, so  V^-1 V0 = L^-1 (W^-1 (P V0)):  a GEMM instead of an inverse.
struct pv0_args { int B, N; const cd* P; const cd* Kx; const cd* Ky; cd* out; };
KH_DEV void pv0_body(const Cta& c, const pv0_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const long long off = (long long)b * n * n;
    const cd* kx = a.Kx + (long long)b * N;
    const cd* ky = a.Ky + (long long)b * N;
    for (int e = c.tid; e < n * N; e += c.nthr) {
        const int i = e / N, g = e - i * N;
        const m22 V0 = v0_block(kx[g], ky[g]);
        const cd p1 = a.P[off + (long long)i * n + g], p2 = a.P[off + (long long)i * n + N + g];
        a.out[off + (long long)i * n + g] = p1 * V0.a + p2 * V0.c;
        a.out[off + (long long)i * n + N + g] = p1 * V0.b + p2 * V0.d;
    }
}
// A = W^-1 + t2 ; B = W^-1 - t2 ; XB = X B ; XA = X A   with t2 = V^-1 V0 given   (alternative.py:182-186; W0 = I)
struct ab2_args { int B, n; const cd* Winv; const cd* t2; const cd* xexp; cd* A; cd* Bm; cd* XB; cd* XA; };
KH_DEV void ab2_body(const Cta& c, const ab2_args& a) {
    const int n = a.n, b = c.bx;
    const long long off = (long long)b * n * n;
    const cd* x = a.xexp + (long long)b * n;
    for (int e = c.tid; e < n * n; e += c.nthr) {
        const int i = e / n;
        const cd w = a.Winv[off + e], t = a.t2[off + e];
        const cd av = w + t, bv = w - t;
        a.A[off + e] = av; a.Bm[off + e] = bv; a.XB[off + e] = x[i] * bv; a.XA[off + e] = x[i] * av;
    }
}

// ------------------------------------------------------------------ flux (crystal.py:363-396, alternative.py:101-128, 235-245)
struct flux_args {
    int B, N;
    const cd* Stot;            // [B][4][n][n]
    const double* wl; const cd* kp; const cd* pol;   // pol [B][2] = (te, tm)
    const double* g;
    cd epsi, epse;
    double* RT;                // [B][2]
    double* orders;            // optional [B][2][N]
};
KH_DEV double cnorm3(cd a, cd b, cd c3) { return sqrt(cabs2(a) + cabs2(b) + cabs2(c3)); }
KH_DEV void flux_body(const Cta& c, const flux_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    double* scratch = (double*)c.smem;
    const double k0 = KH_TWO_PI / a.wl[b];
    const cd kpx = a.kp[2 * b], kpy = a.kp[2 * b + 1];
    // incident(): polarisation vector
    cd te = a.pol[2 * b], tm = a.pol[2 * b + 1];
    double pn = hypot(cabsd(te), cabsd(tm));
    te = (1.0 / pn) * te; tm = (1.0 / pn) * tm;
    cd kzi = cconj(csqrt_(mk(k0 * k0, 0) * a.epsi - kpx * kpx - kpy * kpy));
    double kn = cnorm3(kpx, kpy, kzi);
    cd kbx = (1.0 / kn) * kpx, kby = (1.0 / kn) * kpy, kbz = (1.0 / kn) * kzi;
    cd px, py;
    if (sqrt(cabs2(kpx) + cabs2(kpy)) < 1e-8) { px = te; py = tm; }
    else {
        cd ex = -kby, ey = kbx;                                  // -cross((0,0,-1), kbar)
        double en = sqrt(cabs2(ex) + cabs2(ey));
        ex = (1.0 / en) * ex; ey = (1.0 / en) * ey;
        cd mx = ey * kbz, my = -(ex * kbz), mz = ex * kby - ey * kbx;   // cross(aTE, kbar)
        double mn = cnorm3(mx, my, mz);
        mx = (1.0 / mn) * mx; my = (1.0 / mn) * my;
        px = te * ex + tm * mx; py = te * ey + tm * my;
    }
    const int g0 = (N - 1) / 2;
    const cd kzin = (1.0 / k0) * kzi;                            // poynting_fluxes: kzi / k0
    const cd* S11 = a.Stot + (long long)b * 4 * n * n;
    const cd* S21 = S11 + 2LL * n * n;
    double accR = 0.0, accT = 0.0;
    for (int g = c.tid; g < N; g += c.nthr) {
        cd kxg = mk(kpx.x + a.g[g], kpx.y), kyg = mk(kpy.x + a.g[N + g], kpy.y);
        cd kx = mk(kxg.x / k0, kxg.y / k0), ky = mk(kyg.x / k0, kyg.y / k0);
        for (int side = 0; side < 2; ++side) {
            const cd* S = side ? S21 : S11;
            cd eps = side ? a.epse : a.epsi;
            cd sx = S[(long long)g * n + g0] * px + S[(long long)g * n + N + g0] * py;
            cd sy = S[(long long)(N + g) * n + g0] * px + S[(long long)(N + g) * n + N + g0] * py;
            cd kzf = cconj(csqrt_(mk(k0 * k0, 0) * cconj(eps) - kxg * kxg - kyg * kyg));
            cd kz = mk(kzf.x / k0, kzf.y / k0);
            cd sz = (-(kx * sx + ky * sy)) / kz;
            double t = kz.x / kzin.x * (cabs2(sx) + cabs2(sy) + cabs2(sz));
            if (a.orders) a.orders[((long long)b * 2 + side) * N + g] = t;
            if (side) accT += t; else accR += t;
        }
    }
    accR = cta_sum(c, accR, scratch);
    accT = cta_sum(c, accT, scratch);
    if (c.tid == 0) { a.RT[2 * b] = accR; a.RT[2 * b + 1] = accT; }
}

// ------------------------------------------------------------------ slab S-matrix without an eigensolver
// The reference's legacy solver builds a layer's S-matrix as  T = expm(A dz) on a thin slice, S = matrix_s(T), then
// `slicing_pow` self star products (khepri/tmat/scattering.py:25-51, tmat/matrices.py:167-176, 189-211).  The same idea in
// the Crystal path's field basis (s, u), d/dz' [s; u] = [[0, P], [Q, 0]] [s; u] (alternative.py:158-171), z' = k0 z:
//   exp(x [[0,P],[Q,0]]) = [[I + Om Dc, Sc P], [Q Sc, I + Q Dc P]],   Om = P Q,
//   Sc = x sum_k (x^2 Om)^k / (2k+1)!,   Dc = x^2 sum_k (x^2 Om)^k / (2k+2)!
// i.e. two matrix polynomials in the n x n matrix Om (Paterson-Stockmeyer: powers Om^2..Om^q once, then Horner in Om^q),
// every flop a DMMA GEMM.  dbl_lincomb forms the Horner blocks  sum_i c_i(b) Om^i  with the per-solve slice thickness
// x_b = k0[b] * hx folded into the coefficients.
#define KH_DBL_QMAX 8
struct dbl_lincomb_args {
    int B, n, q;                      // powers I, Om, .., Om^(q-1)
    const cd* pw[KH_DBL_QMAX];        // pw[i] = Om^i as [B][n][n] (pw[0] unused)
    const double* k0; double hx;      // x_b = k0[b] * hx
    double coef[2][KH_DBL_QMAX]; int xpow[2][KH_DBL_QMAX];     // out_o = sum_i coef[o][i] x_b^xpow[o][i] Om^i
    cd* out[2];                       // [B][n][n] each
};
KH_DEV void dbl_lincomb_body(const Cta& c, const dbl_lincomb_args& a) {
    const int n = a.n, b = c.bx, q = a.q;
    const long long off = (long long)b * n * n;
    const double x = a.k0[b] * a.hx;
    double cf[2][KH_DBL_QMAX];
    for (int o = 0; o < 2; ++o)
        for (int i = 0; i < KH_DBL_QMAX; ++i) cf[o][i] = (i < q && a.coef[o][i] != 0.0) ? a.coef[o][i] * pow(x, (double)a.xpow[o][i]) : 0.0;
    const int per = (n * n + 3) / 4, e0 = c.by * per, e1 = (e0 + per < n * n) ? e0 + per : n * n;
    for (int e = e0 + c.tid; e < e1; e += c.nthr) {
        const int i = e / n, j = e - i * n;
        cd v0 = mk(i == j ? cf[0][0] : 0.0, 0.0), v1 = mk(i == j ? cf[1][0] : 0.0, 0.0);
        for (int k = 1; k < q; ++k) { const cd p = a.pw[k][off + e]; v0 = v0 + cf[0][k] * p; v1 = v1 + cf[1][k] * p; }
        a.out[0][off + e] = v0; a.out[1][off + e] = v1;
    }
}

// S-matrix of TWO slices from the transfer matrix M of ONE (thickness h): the slab of thickness 2h is mirror symmetric about its
// mid-plane, so its response splits into an even (u = 0 on the mid-plane) and an odd (s = 0) problem.  With the mode basis of
// the zero-thickness free-space gaps (W0 = I, V0; alternative.py:84-99, fields.py:46-51: s = c+ + c-, u = V0 (c- - c+) at the
// right face), the even fields at the face are [M11; M21] s_mid and the odd ones [M12; M22] u_mid, hence the reflection operators
//   r_e = (M11 - V0^-1 M21) (M11 + V0^-1 M21)^-1,      r_o = (M12 - V0^-1 M22) (M12 + V0^-1 M22)^-1
// and  S11 = S22 = (r_e + r_o) / 2,  S12 = S21 = (r_e - r_o) / 2  (the symmetric form alternative.py:195 returns).  This is
// matrix_s (tmat/matrices.py:167-176) followed by the first multS doubling (tmat/scattering.py:46-49), at the price of two
// inverses and two products instead of one inverse, one product and a full star product.
// dbl_eo forms the four operands  Ee+- = M11 +- V0^-1 M21,  Eo+- = M12 +- V0^-1 M22  (V0^-1: 2x2 blocks of diagonals, O(n^2));
// M11 and M22 arrive without their identity (M11 = I + m11, M22 = I + m22).  Thread <-> (harmonic g, column j): rows g and N + g.
struct dbl_eo_args { int B, N; const cd* m11; const cd* M12; const cd* M21; const cd* m22; const cd* Kx; const cd* Ky;
                     cd* Eep; cd* Eop; cd* Eem; cd* Eom; };
KH_DEV void dbl_eo_body(const Cta& c, const dbl_eo_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const long long off = (long long)b * n * n;
    m22* Vi = (m22*)KH_SMEM(c);                                 // [N]: V0^-1 per harmonic (complex divisions: once per CTA, not per element)
    for (int g = c.tid; g < N; g += c.nthr) Vi[g] = m22_inv(v0_block(a.Kx[(long long)b * N + g], a.Ky[(long long)b * N + g]));
    c.sync();
    const int per = (N * n + 3) / 4, e0 = c.by * per, e1 = (e0 + per < N * n) ? e0 + per : N * n;
    for (int e = e0 + c.tid; e < e1; e += c.nthr) {
        const int g = e / n, j = e - g * n;
        const m22 v = Vi[g];
        const long long r0 = off + (long long)g * n + j, r1 = off + (long long)(N + g) * n + j;
        const cd a0 = a.M21[r0], a1 = a.M21[r1];
        cd b0 = a.m22[r0], b1 = a.m22[r1];
        if (j == g) b0.x += 1.0;
        if (j == N + g) b1.x += 1.0;
        const cd t0 = v.a * a0 + v.b * a1, t1 = v.c * a0 + v.d * a1;          // rows g, N + g of V0^-1 M21
        const cd u0 = v.a * b0 + v.b * b1, u1 = v.c * b0 + v.d * b1;          // rows g, N + g of V0^-1 M22
        cd p0 = a.m11[r0], p1 = a.m11[r1];
        if (j == g) p0.x += 1.0;
        if (j == N + g) p1.x += 1.0;
        const cd q0 = a.M12[r0], q1 = a.M12[r1];
        a.Eep[r0] = p0 + t0; a.Eep[r1] = p1 + t1; a.Eem[r0] = p0 - t0; a.Eem[r1] = p1 - t1;
        a.Eop[r0] = q0 + u0; a.Eop[r1] = q1 + u1; a.Eom[r0] = q0 - u0; a.Eom[r1] = q1 - u1;
    }
}
// S11 = (r_e + r_o) / 2, S12 = (r_e - r_o) / 2
struct dbl_combine_args { int B, n; const cd* re; const cd* ro; MatRef S11, S12; };
KH_DEV void dbl_combine_body(const Cta& c, const dbl_combine_args& a) {
    const int n = a.n, b = c.bx;
    const long long off = (long long)b * n * n;
    cd* o11 = mat_ptr(a.S11, b);
    cd* o12 = mat_ptr(a.S12, b);
    const int per = (n * n + 3) / 4, e0 = c.by * per, e1 = (e0 + per < n * n) ? e0 + per : n * n;
    for (int e = e0 + c.tid; e < e1; e += c.nthr) {
        const int i = e / n, j = e - i * n;
        const cd x = a.re[off + e], y = a.ro[off + e];
        o11[(long long)i * a.S11.ld + j] = 0.5 * (x + y);
        o12[(long long)i * a.S12.ld + j] = 0.5 * (x - y);
    }
}

// All Horner blocks of both series in ONE pass over the powers:  out[j][o] = sum_{i<q} c_o[j q + i](b) Om^i  for j < J.
// (Per Horner step the powers would be read again: q - 1 slabs each time; here they are read once and 2 J slabs are written.)
#define KH_DBL_JMAX 6
struct dbl_blocks_args {
    int B, n, q, J, t;
    const cd* pw[KH_DBL_QMAX];
    const double* k0; double hx;
    cd* out[KH_DBL_JMAX][2];
};
KH_DEV void dbl_blocks_body(const Cta& c, const dbl_blocks_args& a) {
    const int n = a.n, b = c.bx, q = a.q, J = a.J;
    const long long off = (long long)b * n * n;
    double* cf = (double*)KH_SMEM(c);                   // [2][J q]: Sc coefficients x^(2k+1) / (2k+1)!, Dc coefficients x^(2k+2) / (2k+2)!
    if (c.tid == 0) {
        const double x = a.k0[b] * a.hx;
        double ps = x, pd = 0.5 * x * x;                // k = 0
        for (int k = 0; k < J * q; ++k) {
            cf[k] = (k < a.t) ? ps : 0.0;
            cf[J * q + k] = (k < a.t - 1) ? pd : 0.0;
            ps *= x * x / ((2.0 * k + 2.0) * (2.0 * k + 3.0));
            pd *= x * x / ((2.0 * k + 3.0) * (2.0 * k + 4.0));
        }
    }
    c.sync();
    const int per = (n * n + 3) / 4, e0 = c.by * per, e1 = (e0 + per < n * n) ? e0 + per : n * n;
    for (int e = e0 + c.tid; e < e1; e += c.nthr) {
        const int i = e / n, j2 = e - i * n;
        cd p[KH_DBL_QMAX];
        p[0] = mk(i == j2 ? 1.0 : 0.0, 0.0);
#pragma unroll
        for (int k = 1; k < KH_DBL_QMAX; ++k) p[k] = (k < q) ? a.pw[k][off + e] : mk(0.0, 0.0);
        for (int j = 0; j < J; ++j) {
            cd v0 = mk(0, 0), v1 = mk(0, 0);
#pragma unroll
            for (int k = 0; k < KH_DBL_QMAX; ++k) if (k < q) { v0 = v0 + cf[j * q + k] * p[k]; v1 = v1 + cf[J * q + j * q + k] * p[k]; }
            a.out[j][0][off + e] = v0; a.out[j][1][off + e] = v1;
        }
    }
}

// safety net for the (slices, terms) the host chose from its bound on the spectrum: theta_b = x_b sqrt(||Om||_1) must stay
// below theta_lim, otherwise bit 2 of info is raised for that solve (the truncated series may have lost accuracy).
struct dbl_check_args { int B, n; const cd* Om; const double* k0; double hx, theta_lim; int* info; };
KH_DEV void dbl_check_body(const Cta& c, const dbl_check_args& a) {
    const int n = a.n, b = c.bx;
    const cd* A = a.Om + (long long)b * n * n;
    double* scratch = (double*)c.smem;
    double m = 0.0;
    for (int j = c.tid; j < n; j += c.nthr) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += cabsd(A[(long long)i * n + j]);
        m = fmax(m, s);
    }
    m = cta_max(c, m, scratch);
    const double th = a.k0[b] * a.hx * sqrt(m);
    if (c.tid == 0 && !(th <= a.theta_lim)) KH_ATOMIC_OR(&a.info[b], 4);
}

// Conditioning guard of a self star product  Y = D^-1 S12,  D = I - S11^2.  At a resonance of the sub-slab the doubling method
// passes through (a guided mode of a slab of depth d / 2^k between vacuum gaps) D is nearly singular and the doubled S-matrix
// loses digits that the eigen-decomposition does not.  The relative error of the computed Y is bounded by
//     eps * kappa(D) * c,    c = ||D^-1|| ||S12|| / ||Y||   (the cancellation in the product with the explicit inverse),
// measured in 1-norms in two passes (after the inverse: kappa(D) ||D^-1|| -> scratch; after the product: times ||S12|| / ||Y||).
// Above the limit info bit 3 is raised and the host re-solves that source with the eigen method (Engine.solve_batch, "auto").
struct dbl_cond_args { int B, n, mode; MatRef X, Z; double* scratch; double limit; int* info; };
KH_DEV void dbl_cond_body(const Cta& c, const dbl_cond_args& a) {        // (1-norms: column sums; lane <-> column, a warp per group of 32 columns)
    const int n = a.n, b = c.bx;
    const cd* X = mat_ptr(a.X, b);
    const cd* Z = mat_ptr(a.Z, b);
    double* red = (double*)c.smem;
    double m1 = 0.0, m2 = 0.0;
#ifdef KH_HOST_EMU
    for (int j = 0; j < n; ++j) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < n; ++i) { s1 += cabsd(X[(long long)i * a.X.ld + j]); s2 += cabsd(Z[(long long)i * a.Z.ld + j]); }
        m1 = fmax(m1, s1); m2 = fmax(m2, s2);
    }
#else
    const int lane = c.tid & 31, warp = c.tid >> 5, nw = c.nthr >> 5;
    for (int j0 = warp * 32; j0 < n; j0 += nw * 32) {
        const int j = j0 + lane;
        if (j < n) {
            double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
            int i = 0;
            for (; i + 4 <= n; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { s1[u] += cabsd(X[(long long)(i + u) * a.X.ld + j]); s2[u] += cabsd(Z[(long long)(i + u) * a.Z.ld + j]); }
            }
            for (; i < n; ++i) { s1[0] += cabsd(X[(long long)i * a.X.ld + j]); s2[0] += cabsd(Z[(long long)i * a.Z.ld + j]); }
            m1 = fmax(m1, (s1[0] + s1[1]) + (s1[2] + s1[3])); m2 = fmax(m2, (s2[0] + s2[1]) + (s2[2] + s2[3]));
        }
    }
#endif
    m1 = cta_max(c, m1, red);
    m2 = cta_max(c, m2, red);
    if (c.tid == 0) {
        if (a.mode == 0) a.scratch[b] = m1 * m2 * m2;                  // X = D, Z = D^-1
        else if (!(a.scratch[b] * m1 / m2 <= a.limit)) KH_ATOMIC_OR(&a.info[b], 8);      // X = S12, Z = Y
    }
}

// ------------------------------------------------------------------ two columns of a product (flux columns of the star chain)
// Cout[:, c] = A B[:, c] (+ Cin[:, c]) for c in {c0, c1}: matrix-vector work, bound by reading A once (n^2 16 B per solve).
// One CTA per solve; the two B columns are staged in shared memory, warp <-> row, lanes stride over k, shuffle reduction.
struct zgemv2_args { int n, c0, c1; MatRef A, B, Cin, Cout; };
KH_DEV void zgemv2_body(const Cta& c, const zgemv2_args& a) {
    const int n = a.n, b = c.bx;
    const cd* A = mat_ptr(a.A, b);
    const cd* Bm = mat_ptr(a.B, b);
    const cd* Cin = mat_ptr(a.Cin, b);
    cd* Co = mat_ptr(a.Cout, b);
    cd* x0 = (cd*)KH_SMEM(c);
    cd* x1 = x0 + n;
    for (int k = c.tid; k < n; k += c.nthr) { x0[k] = Bm[(long long)k * a.B.ld + a.c0]; x1[k] = Bm[(long long)k * a.B.ld + a.c1]; }
    c.sync();
    const int lane = c.tid % KH_WARP, warp = c.tid / KH_WARP, nw = (c.nthr + KH_WARP - 1) / KH_WARP;
    for (int i = warp; i < n; i += nw) {
        const cd* ar = A + (long long)i * a.A.ld;
        cd s0 = mk(0, 0), s1 = mk(0, 0);
        for (int k = lane; k < n; k += KH_WARP) { const cd v = ar[k]; cfma(s0, v, x0[k]); cfma(s1, v, x1[k]); }
        s0.x = kh_warp_allsum(s0.x); s0.y = kh_warp_allsum(s0.y); s1.x = kh_warp_allsum(s1.x); s1.y = kh_warp_allsum(s1.y);
        if (lane == 0) {
            if (Cin) { s0 = s0 + Cin[(long long)i * a.Cin.ld + a.c0]; s1 = s1 + Cin[(long long)i * a.Cin.ld + a.c1]; }
            Co[(long long)i * a.Cout.ld + a.c0] = s0; Co[(long long)i * a.Cout.ld + a.c1] = s1;
        }
    }
}
