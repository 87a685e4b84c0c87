// Convolution matrix of a pixmap layer (khepri/tools.py:33-56).
//
// The reference takes a full fft2 of the Nx x Ny pixmap and then gathers (2P-1)(2Q-1) of its
// coefficients into an N x N Toeplitz-block matrix with a Python double loop, every solve.  Only
// those few coefficients are needed, so this is a *pruned* separable DFT (HBM-bound: the pixmap is
// read exactly once, coalesced along y) followed by a pure-index gather (bit exact):
//   stage 1:  G[x][l]   = sum_y eps[x][y] * w_Ny^(l*y)          l = -(Q-1)..(Q-1)
//   stage 2:  F[m][l]   = sum_x G[x][l] * w_Nx^(m*x) / (Nx*Ny)  m = -(P-1)..(P-1)
//   gather :  C[qr*P+pr][qc*P+pc] = F[pr-pc][qr-qc]
// Twiddles come from an exact integer reduction (l*y mod Ny) and sincospi, so the coefficients
// agree with pocketfft to ~1e-16 relative to max|F|.
#pragma once
#include "kh_common.cuh"

struct dft1_args {
    int Nx, Ny, Q, is_complex;
    const void* pix;        // [L][Nx][Ny] f64 or c128
    cd* G;                  // [L][Nx][2Q-1]
};
KH_DEV cd twiddle(long long num, int den) {   // exp(-2 pi i num / den), num already reduced mod den
    double s, c;
    double t = -2.0 * (double)num / (double)den;
#ifdef __CUDA_ARCH__
    sincospi(t, &s, &c);
#else
    s = sin(M_PI * t); c = cos(M_PI * t);
    // exact values at the quarter points keep the emulation as clean as sincospi
    if (4 * num == (long long)den) { s = -1.0; c = 0.0; }
    else if (2 * num == (long long)den) { s = 0.0; c = -1.0; }
    else if (4 * num == 3LL * den) { s = 1.0; c = 0.0; }
    else if (num == 0) { s = 0.0; c = 1.0; }
#endif
    return mk(c, s);
}
KH_DEV void dft1_body(const Cta& c, const dft1_args& a) {
    const int x = c.bx, lay = c.by, Ny = a.Ny, nl = 2 * a.Q - 1;
    cd* tw = (cd*)c.smem;                      // [Ny]
    cd* row = tw + Ny;                         // [Ny]
    for (int j = c.tid; j < Ny; j += c.nthr) {
        tw[j] = twiddle(j, Ny);
        long long off = ((long long)lay * a.Nx + x) * Ny + j;
        row[j] = a.is_complex ? ((const cd*)a.pix)[off] : mk(((const double*)a.pix)[off], 0.0);
    }
    c.sync();
    // twiddle index (l*y) mod Ny advances by a constant step: no integer division in the inner loop
    const int wlanes = KH_WARP;
#ifdef KH_HOST_EMU
    const int warp = 0, lane = 0, nw = 1;
#else
    const int warp = c.tid >> 5, lane = c.tid & 31, nw = c.nthr >> 5;
#endif
    for (int li = warp; li < nl; li += nw) {
        const int l = li - (a.Q - 1);
        const int lm = ((l % Ny) + Ny) % Ny;                                  // l mod Ny in [0, Ny)
        int idx = (int)(((long long)lm * lane) % Ny);
        const int step = (int)(((long long)lm * wlanes) % Ny);
        cd acc0 = mk(0, 0), acc1 = mk(0, 0);
        int y = lane;
        for (; y + wlanes < Ny; y += 2 * wlanes) {
            int idx2 = idx + step; if (idx2 >= Ny) idx2 -= Ny;
            cfma(acc0, row[y], tw[idx]);
            cfma(acc1, row[y + wlanes], tw[idx2]);
            idx = idx2 + step; if (idx >= Ny) idx -= Ny;
        }
        if (y < Ny) cfma(acc0, row[y], tw[idx]);
        cd acc = acc0 + acc1;
        acc.x = kh_warp_allsum(acc.x); acc.y = kh_warp_allsum(acc.y);
        if (lane == 0) a.G[((long long)lay * a.Nx + x) * nl + li] = acc;
    }
}

// Stage 1, row-per-thread form (the one the launcher uses for 2Q-1 <= 29): thread <-> pixmap row x, all 2Q-1 running sums
// of the row in registers.  The pixmap streams through shared memory in coalesced [64 rows][32 columns] tiles (read once,
// 256-byte segments), the twiddle of (l, y) is the same for every row, i.e. ONE broadcast LDS per 64 rows, and there is no
// cross-lane reduction at all -- the warp-per-coefficient form above spends more instructions on shuffles than on FMAs.
// Per pixel: 2 (real pixmap) or 4 (complex) DFMA per coefficient; at 29 coefficients that is 14.5 flop per byte of pixmap,
// above the B200's FP64 balance of 5.7 flop/B: the stage is FP64-bound, not HBM-bound, for the larger bases.
#define DFT1R_ROWS 64
#define DFT1R_YT 32
template <int NL, bool CPLX>
KH_DEV void dft1_rows_body_t(const Cta& c, const dft1_args& a) {
    const int Ny = a.Ny, nl = 2 * a.Q - 1, lay = c.by, x0 = c.bx * DFT1R_ROWS;
    cd* tw = (cd*)c.smem;                                           // [Ny]
    double* tile = (double*)(tw + Ny);                              // [ROWS][pitch] (re, or re/im interleaved)
    constexpr int W = CPLX ? 2 : 1, PITCH = DFT1R_YT * W + 1;
    for (int j = c.tid; j < Ny; j += c.nthr) tw[j] = twiddle(j, Ny);
    cd acc[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) acc[l] = mk(0.0, 0.0);
    const int row = c.tid;                                          // this thread's row inside the CTA (nthr == ROWS; emulation: loop)
    const double* pix = (const double*)a.pix + ((long long)lay * a.Nx + x0) * Ny * W;
    const int rows = (a.Nx - x0 < DFT1R_ROWS) ? a.Nx - x0 : DFT1R_ROWS;
#ifdef KH_HOST_EMU
    for (int r = 0; r < rows; ++r) {                                // one virtual thread: plain loops, same arithmetic order
        cd accr[NL];
        for (int l = 0; l < NL; ++l) accr[l] = mk(0.0, 0.0);
        for (int y = 0; y < Ny; ++y) {
            const double re = pix[((long long)r * Ny + y) * W], im = CPLX ? pix[((long long)r * Ny + y) * W + 1] : 0.0;
            int idx = (int)((((long long)(-(a.Q - 1)) * y) % Ny + Ny) % Ny);
            for (int l = 0; l < NL; ++l) {
                if (l < nl) { const cd t = tw[idx]; if (CPLX) cfma(accr[l], mk(re, im), t); else { accr[l].x = fma(re, t.x, accr[l].x); accr[l].y = fma(re, t.y, accr[l].y); } }
                idx += y % Ny; if (idx >= Ny) idx -= Ny;
            }
        }
        for (int l = 0; l < nl; ++l) a.G[((long long)lay * a.Nx + x0 + r) * nl + l] = accr[l];
    }
    (void)row; (void)tile; (void)acc;
#else
    for (int y0 = 0; y0 < Ny; y0 += DFT1R_YT) {
        const int yt = (Ny - y0 < DFT1R_YT) ? Ny - y0 : DFT1R_YT;
        c.sync();                                                   // previous tile consumed (and twiddles written, first pass)
        {   // nthr == ROWS == 64: a warp reads 256-byte row segments; 8 independent loads in flight per thread
            constexpr int SEG = DFT1R_YT * W, PER = DFT1R_ROWS * SEG / DFT1R_ROWS;          // elements per thread per tile
            const int cc = c.tid % SEG, rb = c.tid / SEG, rstep = DFT1R_ROWS / SEG;          // (SEG = 32 or 64 divides / equals the CTA)
            const bool cok = cc < yt * W;
#pragma unroll
            for (int i0 = 0; i0 < PER; i0 += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int r = rb + (i0 + u) * rstep; v[u] = (r < rows && cok) ? pix[(long long)r * Ny * W + (long long)y0 * W + cc] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int r = rb + (i0 + u) * rstep; tile[r * PITCH + cc] = v[u]; }
            }
        }
        c.sync();
        // twiddle index of (l, y) = ((l - (Q-1)) y) mod Ny, walked as FOUR interleaved chains (l mod 4) so that the dependent
        // integer chain per pixel is NL/4 steps of depth 2 (add, unsigned min) instead of NL steps; no predicate on l in the
        // loop: NL >= nl and the surplus sums are simply not stored
        int ystep = y0 % Ny;
        int idx0 = (int)((((long long)(-(a.Q - 1)) * y0) % Ny + Ny) % Ny);
        const int dl0 = (int)((((long long)(-(a.Q - 1))) % Ny + Ny) % Ny);
        for (int yy = 0; yy < yt; ++yy) {
            const double re = tile[row * PITCH + yy * W];
            const double im = CPLX ? tile[row * PITCH + yy * W + 1] : 0.0;
            unsigned ic[4];
            ic[0] = (unsigned)idx0;
#pragma unroll
            for (int r = 1; r < 4; ++r) { const unsigned t = ic[r - 1] + (unsigned)ystep; ic[r] = min(t, t - (unsigned)Ny); }
            unsigned st4 = (unsigned)ystep * 4u; st4 = st4 % (unsigned)Ny;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const cd t = tw[ic[l & 3]];
                if (CPLX) cfma(acc[l], mk(re, im), t);
                else { acc[l].x = fma(re, t.x, acc[l].x); acc[l].y = fma(re, t.y, acc[l].y); }
                const unsigned nx = ic[l & 3] + st4; ic[l & 3] = min(nx, nx - (unsigned)Ny);
            }
            idx0 += dl0; if (idx0 >= Ny) idx0 -= Ny;
            ++ystep; if (ystep >= Ny) ystep -= Ny;
        }
    }
    if (row < rows) {
#pragma unroll
        for (int l = 0; l < NL; ++l) if (l < nl) a.G[((long long)lay * a.Nx + x0 + row) * nl + l] = acc[l];
    }
#endif
}
#define DFT1R_INST(NL) \
    KH_DEV void dft1_rows##NL##r_body(const Cta& c, const dft1_args& a) { dft1_rows_body_t<NL, false>(c, a); } \
    KH_DEV void dft1_rows##NL##c_body(const Cta& c, const dft1_args& a) { dft1_rows_body_t<NL, true>(c, a); }
DFT1R_INST(5) DFT1R_INST(9) DFT1R_INST(13) DFT1R_INST(17) DFT1R_INST(21) DFT1R_INST(25) DFT1R_INST(29)
static inline size_t dft1_rows_smem(int Ny, int is_complex) { return (size_t)Ny * sizeof(cd) + (size_t)DFT1R_ROWS * (DFT1R_YT * (is_complex ? 2 : 1) + 1) * sizeof(double) + 16; }
static inline int dft1_launch(kh_stream_t st, int L, const dft1_args& a) {
    const int nl = 2 * a.Q - 1;
    const dim3 grid((a.Nx + DFT1R_ROWS - 1) / DFT1R_ROWS, L);
    const size_t sm = dft1_rows_smem(a.Ny, a.is_complex);
    if (nl <= 29 && sm <= (size_t)200 * 1024) {
#define DFT1R_GO(NL) if (nl <= NL) return a.is_complex ? kh_launch<dft1_args, dft1_rows##NL##c_body>(grid, DFT1R_ROWS, sm, st, a, "dft1") \
                                                       : kh_launch<dft1_args, dft1_rows##NL##r_body>(grid, DFT1R_ROWS, sm, st, a, "dft1");
        DFT1R_GO(5) DFT1R_GO(9) DFT1R_GO(13) DFT1R_GO(17) DFT1R_GO(21) DFT1R_GO(25) DFT1R_GO(29)
#undef DFT1R_GO
    }
    return kh_launch<dft1_args, dft1_body>(dim3(a.Nx, L), 256, (size_t)2 * a.Ny * sizeof(cd), st, a, "dft1");
}

struct dft2_args {
    int Nx, Ny, P, Q;
    const cd* G;            // [L][Nx][2Q-1]
    cd* F;                  // [L][2P-1][2Q-1]
};
KH_DEV void dft2_body(const Cta& c, const dft2_args& a) {
    const int nl = 2 * a.Q - 1, nm = 2 * a.P - 1, lay = c.by;
    const int mi = c.bx / nl, li = c.bx % nl;
    const int m = mi - (a.P - 1);
    double* scratch = (double*)c.smem;
    cd acc = mk(0, 0);
    for (int x = c.tid; x < a.Nx; x += c.nthr) {
        long long idx = (((long long)m * x) % a.Nx + a.Nx) % a.Nx;
        cfma(acc, a.G[((long long)lay * a.Nx + x) * nl + li], twiddle(idx, a.Nx));
    }
    double re = cta_sum(c, acc.x, scratch);
    double im = cta_sum(c, acc.y, scratch);
    if (c.tid == 0) {
        double cnt = (double)a.Nx * (double)a.Ny;
        a.F[((long long)lay * nm + mi) * nl + li] = mk(re / cnt, im / cnt);
    }
}

#ifndef KH_HOST_EMU
// Stage 2, table form: one CTA per (layer, m); lane <-> l reads G[x][0..nl) as one contiguous segment per x, the four
// warps split the x range, the twiddle exp(-2 pi i m x / Nx) comes from a per-CTA table walked with a running index
// (the form above evaluates sincospi once per term).
__device__ __forceinline__ void dft2_rows_body(const Cta& c, const dft2_args& a) {
    const int nl = 2 * a.Q - 1, nm = 2 * a.P - 1, lay = c.by, mi = c.bx, m = mi - (a.P - 1), Nx = a.Nx;
    cd* tw = (cd*)c.smem;                     // [Nx]
    cd* part = tw + Nx;                       // [4][32]
    for (int j = c.tid; j < Nx; j += c.nthr) tw[j] = twiddle(j, Nx);
    c.sync();
    const int warp = c.tid >> 5, lane = c.tid & 31, nw = c.nthr >> 5;
    const int xb = (int)((long long)Nx * warp / nw), xe = (int)((long long)Nx * (warp + 1) / nw);
    const unsigned mm = (unsigned)(((m % Nx) + Nx) % Nx);
    unsigned idx = (unsigned)(((long long)mm * xb) % Nx);
    cd acc0 = mk(0.0, 0.0), acc1 = mk(0.0, 0.0);
    const cd* g = a.G + ((long long)lay * Nx + xb) * nl + lane;
    const bool on = lane < nl;
    int x = xb;
    for (; x + 1 < xe; x += 2, g += 2 * nl) {
        const cd t0 = tw[idx]; unsigned i1 = idx + mm; i1 = min(i1, i1 - (unsigned)Nx);
        const cd t1 = tw[i1]; idx = i1 + mm; idx = min(idx, idx - (unsigned)Nx);
        if (on) { cfma(acc0, g[0], t0); cfma(acc1, g[nl], t1); }
    }
    if (x < xe && on) cfma(acc0, g[0], tw[idx]);
    part[warp * 32 + lane] = acc0 + acc1;
    c.sync();
    if (warp == 0 && on) {
        cd s = part[lane];
        for (int w = 1; w < nw; ++w) s = s + part[w * 32 + lane];
        const double cnt = (double)a.Nx * (double)a.Ny;
        a.F[((long long)lay * nm + mi) * nl + lane] = mk(s.x / cnt, s.y / cnt);
    }
}
#endif

// Gather from the compact coefficient table F[2P-1][2Q-1] (centre = zero frequency)
struct gather_args { int P, Q; const cd* F; cd* C; };
KH_DEV void gather_body(const Cta& c, const gather_args& a) {
    const int N = a.P * a.Q, nl = 2 * a.Q - 1, nm = 2 * a.P - 1, lay = c.by;
    const cd* F = a.F + (long long)lay * nm * nl;
    cd* C = a.C + (long long)lay * N * N;
    for (int e = c.bx * c.nthr + c.tid; e < N * N; e += c.nthr * 64) {   // grid.x = 64 CTAs
        int row = e / N, col = e - row * N;
        int pr = row % a.P, qr = row / a.P, pc = col % a.P, qc = col / a.P;
        C[e] = F[(long long)(pr - pc + a.P - 1) * nl + (qr - qc + a.Q - 1)];
    }
}

// Gather from a full, already shifted coefficient array (tools.convolution_matrix_fourier, bit exact,
// python negative-index wrap included)
struct gather_full_args { int P, Q, Nx, Ny; const cd* F; cd* C; int* err; };
KH_DEV void gather_full_body(const Cta& c, const gather_full_args& a) {
    const int N = a.P * a.Q;
    for (int e = c.bx * c.nthr + c.tid; e < N * N; e += c.nthr * 64) {
        int row = e / N, col = e - row * N;
        int pr = row % a.P, qr = row / a.P, pc = col % a.P, qc = col / a.P;
        int ix = a.Nx / 2 + (pr - pc), iy = a.Ny / 2 + (qr - qc);
        if (ix < 0) ix += a.Nx;
        if (iy < 0) iy += a.Ny;
        if (ix < 0 || ix >= a.Nx || iy < 0 || iy >= a.Ny) { *a.err = 1; a.C[e] = mk(0, 0); }
        else a.C[e] = a.F[(long long)ix * a.Ny + iy];
    }
}
