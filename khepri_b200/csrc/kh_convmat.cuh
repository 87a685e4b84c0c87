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
