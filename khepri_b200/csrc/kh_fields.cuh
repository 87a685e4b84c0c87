// Field reconstruction (khepri/crystal.py:234-343, khepri/fields.py, khepri/fourier.py:136-142).
//
// The reference redoes, for every z-slice, a 4N x 4N LU of the layer eigenbasis R = [[W, W], [-V, V]],
// the mode-amplitude translation and a dense N x (nx*ny) phase matrix.  Everything that does not
// depend on z is hoisted here:
//   per (solve, stack position i):  c+ = (I - Sl22 Sr11)^-1 Sl21 c1p ; c- = Sr11 c+          (fields.py:18-27)
//                                   m  = R_i^-1 R_0 [c+; c-]  using the block inverse
//                                        R^-1 = 1/2 [[W^-1, -V^-1], [W^-1, V^-1]]
//   per (solve, z):                 t = exp(+-lambda k0 (d - zr)) (.) m   (clipped at 1e14)      (fields.py:29-31,53-62)
//                                   (sx,sy) = W (t1 + t2) ; (ux,uy) = V (t2 - t1) ; sz, uz      (fields.py:68-76)
//   per solve:                      one DMMA GEMM  [6 nz x N] . [N x npts]  with the phase matrix
//                                   exp(i (kx_g x_p + ky_g y_p))                                 (fourier.py:136-142)
// Included at the end of kh_api.cu (needs kh_plan and the launch helpers).
#pragma once
#define KH_FIELDS_IMPL 1

// y[i] = sum_j M[i][j] x[j]  (all threads of the CTA take part; caller synchronises afterwards)
KH_DEV void cta_matvec(const Cta& c, const cd* M, long long ld, const cd* x, cd* y, int rows, int cols) {
#ifdef KH_HOST_EMU
    (void)c;
    for (int i = 0; i < rows; ++i) {
        cd acc = mk(0, 0);
        for (int j = 0; j < cols; ++j) cfma(acc, M[i * ld + j], x[j]);
        y[i] = acc;
    }
#else
    const int warp = c.tid >> 5, lane = c.tid & 31, nw = c.nthr >> 5;
    for (int i = warp; i < rows; i += nw) {
        cd acc = mk(0, 0);
        for (int j = lane; j < cols; j += 32) cfma(acc, M[i * ld + j], x[j]);
        for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
        if (lane == 0) y[i] = acc;
    }
#endif
}

// c1p = first half of R_ref^-1 [e; h] = 1/2 (e - V_ref^-1 h)          (crystal.py:259-262, W_ref = I)
struct fld_c1p_args { int B, N; cd eps; const cd* Kx; const cd* Ky; const cd* inc; cd* c1p; };
KH_DEV void fld_c1p_body(const Cta& c, const fld_c1p_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const cd* e = a.inc + (long long)b * 2 * n;
    const cd* h = e + n;
    for (int g = c.tid; g < N; g += c.nthr) {
        cd kx = a.Kx[(long long)b * N + g], ky = a.Ky[(long long)b * N + g];
        cd lam = times_i(cconj(csqrt_(a.eps - kx * kx - ky * ky)));
        m22 Vi = m22_inv(q_over_lam(kx, ky, a.eps, lam));
        cd v1 = Vi.a * h[g] + Vi.b * h[N + g], v2 = Vi.c * h[g] + Vi.d * h[N + g];
        a.c1p[(long long)b * n + g] = 0.5 * (e[g] - v1);
        a.c1p[(long long)b * n + N + g] = 0.5 * (e[N + g] - v2);
    }
}

// amplitudes in the gap right of stack position i and their free-space field vector y = R0 [c+; c-]
// (one CTA per (solve b, stack position of the run): bx = b * run + j, which is also the batch index of the three matrix stacks)
struct fld_amp_args {
    int B, N, run;
    MatRef Finv, Sl21, Sr11;
    const cd* c1p; const cd* Kx; const cd* Ky;
    cd* y12;                       // [B * run][2][n]
};
KH_DEV void fld_amp_body(const Cta& c, const fld_amp_args& a) {
    const int N = a.N, n = 2 * N, bj = c.bx, b = bj / a.run;
    cd* t0 = (cd*)c.smem; cd* cp = t0 + n; cd* cm = cp + n;
    cta_matvec(c, mat_ptr(a.Sl21, bj), a.Sl21.ld, a.c1p + (long long)b * n, t0, n, n);
    c.sync();
    cta_matvec(c, mat_ptr(a.Finv, bj), a.Finv.ld, t0, cp, n, n);
    c.sync();
    cta_matvec(c, mat_ptr(a.Sr11, bj), a.Sr11.ld, cp, cm, n, n);
    c.sync();
    cd* y1 = a.y12 + (long long)bj * 2 * n; cd* y2 = y1 + n;
    for (int g = c.tid; g < N; g += c.nthr) {
        m22 V0 = v0_block(a.Kx[(long long)b * N + g], a.Ky[(long long)b * N + g]);
        cd d1 = cm[g] - cp[g], d2 = cm[N + g] - cp[N + g];
        y1[g] = cp[g] + cm[g]; y1[N + g] = cp[N + g] + cm[N + g];
        y2[g] = V0.a * d1 + V0.b * d2; y2[N + g] = V0.c * d1 + V0.d * d2;
    }
}

// m = R_i^-1 y :  m1 = 1/2 (W^-1 y1 - V^-1 y2),  m2 = 1/2 (W^-1 y1 + V^-1 y2)
// Winv, Vinv: [B][nL][n][n]; the layer of stack position i0 + j from the device copy of the stack
struct fld_modes_args { int B, n, run, i0, nL; const int* stack; const cd* Winv; const cd* Vinv; const cd* y12; cd* m12; long long m_bstride; };
KH_DEV void fld_modes_body(const Cta& c, const fld_modes_args& a) {
    const int n = a.n, bj = c.bx, b = bj / a.run, i = a.i0 + (bj - b * a.run);
    cd* p = (cd*)c.smem; cd* q = p + n;
    const cd* y1 = a.y12 + (long long)bj * 2 * n;
    const long long wv = ((long long)b * a.nL + a.stack[i]) * n * n;
    cta_matvec(c, a.Winv + wv, n, y1, p, n, n);
    cta_matvec(c, a.Vinv + wv, n, y1 + n, q, n, n);
    c.sync();
    cd* m1 = a.m12 + (long long)b * a.m_bstride + (long long)i * 2 * n; cd* m2 = m1 + n;
    for (int i = c.tid; i < n; i += c.nthr) { m1[i] = 0.5 * (p[i] - q[i]); m2[i] = 0.5 * (p[i] + q[i]); }
}

// Fourier field vectors at one depth: out[b][z][6][N] = (sx, sy, sz, ux, uy, uz)
struct fld_z_args {
    int B, N, nz;
    const int* zpos;               // [nz] stack position of each depth
    const int* zlayer;             // [nz] layer-table index of each depth
    const double* zdist;           // [nz] d_layer - z_relative
    const cd* m12; long long m_bstride, m_pstride;     // [B][Ls][2][n]
    const cd* W; const cd* V; const cd* L; long long wv_bstride; int nL;   // [B][nL][n][n], [B][nL][n]
    const cd* const* IC;           // [nL] device pointers to N x N inverse convolution matrices (null: scalar)
    const cd* ICs;                 // [nL] scalar 1/eps for the layers without a matrix
    const cd* Kx; const cd* Ky; const double* k0;
    cd* out;
};
KH_DEV void fld_z_body(const Cta& c, const fld_z_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx, iz = c.by;
    const int li = a.zlayer[iz];
    cd* ts = (cd*)c.smem; cd* td = ts + n; cd* sxy = td + n; cd* uxy = sxy + n; cd* rhs = uxy + n;
    const cd* m1 = a.m12 + (long long)b * a.m_bstride + (long long)a.zpos[iz] * a.m_pstride;
    const cd* m2 = m1 + n;
    const cd* lam = a.L + ((long long)b * a.nL + li) * n;
    const double zb = a.k0[b] * a.zdist[iz];
    for (int i = c.tid; i < n; i += c.nthr) {
        cd e1 = cexp_(zb * lam[i]), e2 = cexp_((-zb) * lam[i]);
        double a1 = cabsd(e1), a2 = cabsd(e2);
        if (a1 > 1e14) e1 = (1.0 / (a1 / 1e14)) * e1;            // fields.py:58-60
        if (a2 > 1e14) e2 = (1.0 / (a2 / 1e14)) * e2;
        cd t1 = e1 * m1[i], t2 = e2 * m2[i];
        ts[i] = t1 + t2; td[i] = t2 - t1;
    }
    c.sync();
    const cd* W = a.W + ((long long)b * a.nL + li) * n * n;
    const cd* V = a.V + ((long long)b * a.nL + li) * n * n;
    cta_matvec(c, W, n, ts, sxy, n, n);
    cta_matvec(c, V, n, td, uxy, n, n);
    c.sync();
    const cd* kx = a.Kx + (long long)b * N; const cd* ky = a.Ky + (long long)b * N;
    cd* o = a.out + ((long long)b * a.nz + iz) * 6 * N;
    for (int g = c.tid; g < N; g += c.nthr) {
        cd sx = sxy[g], sy = sxy[N + g], ux = uxy[g], uy = uxy[N + g];
        o[g] = sx; o[N + g] = sy; o[3 * N + g] = ux; o[4 * N + g] = uy;
        cd w = kx[g] * sy - ky[g] * sx;
        o[5 * N + g] = mk(w.y, -w.x);                            // uz = -i (kx sy - ky sx)
        rhs[g] = kx[g] * uy - ky[g] * ux;
    }
    c.sync();
    const cd* ICm = a.IC[li];
    if (ICm) {
        cta_matvec(c, ICm, N, rhs, ts, N, N);
        c.sync();
        for (int g = c.tid; g < N; g += c.nthr) o[2 * N + g] = mk(ts[g].y, -ts[g].x);   // sz = -i IC rhs
    } else {
        cd s = a.ICs[li];
        for (int g = c.tid; g < N; g += c.nthr) { cd w = s * rhs[g]; o[2 * N + g] = mk(w.y, -w.x); }   // -1j * IC * rhs
    }
}

// phase matrix Ph[b][g][p] = exp(i (kx_g x_p + ky_g y_p)),  kx_g = kp_x + g_x (not normalised)
struct fld_phase_args { int B, N, npts; const cd* kp; const double* g; const double* x; const double* y; cd* Ph; double sign = 1.0; };
KH_DEV void fld_phase_body(const Cta& c, const fld_phase_args& a) {
    const int b = c.bx, g = c.by;
    cd kx = mk(a.kp[2 * b].x + a.g[g], a.kp[2 * b].y), ky = mk(a.kp[2 * b + 1].x + a.g[a.N + g], a.kp[2 * b + 1].y);
    cd* o = a.Ph + ((long long)b * a.N + g) * a.npts;
    for (int p = c.tid; p < a.npts; p += c.nthr) {
        cd arg = a.sign * (a.x[p] * kx + a.y[p] * ky);
        o[p] = cexp_(mk(-arg.y, arg.x));
    }
}

// ---- separable inverse transform on a rectangular (x, y) grid (fields_volume with meshgrid coordinates) for
// lattices whose ky depends on the q index only (b1 along x: square / rectangular lattices):
//     f[iy][ix] = sum_q ( sum_p S[qP+p] exp(i kx_{pq} x_ix) ) exp(i ky_q y_iy)
// 8 (N nx + Q nx ny) flops per map instead of the 8 N nx ny of the dense phase matrix, which moves the field maps from the
// FP64 pipe to the HBM roofline (16 B written per Q complex FMAs).  Tables: Xt[b][g][ix], Yt[b][q][iy].
struct fld_gtab_args { int B, N, P, Q, nx, ny; const cd* kp; const double* g; const double* xs; const double* ys; cd* Xt; cd* Yt; };
KH_DEV void fld_gtab_body(const Cta& c, const fld_gtab_args& a) {
    const int b = c.bx, r = c.by;                     // r < N: row g of Xt ; r >= N: row q of Yt
    if (r < a.N) {
        const cd kx = mk(a.kp[2 * b].x + a.g[r], a.kp[2 * b].y);
        cd* o = a.Xt + ((long long)b * a.N + r) * a.nx;
        for (int i = c.tid; i < a.nx; i += c.nthr) { const cd arg = a.xs[i] * kx; o[i] = cexp_(mk(-arg.y, arg.x)); }
    } else {
        const int q = r - a.N;
        const cd ky = mk(a.kp[2 * b + 1].x + a.g[a.N + q * a.P], a.kp[2 * b + 1].y);
        cd* o = a.Yt + ((long long)b * a.Q + q) * a.ny;
        for (int i = c.tid; i < a.ny; i += c.nthr) { const cd arg = a.ys[i] * ky; o[i] = cexp_(mk(-arg.y, arg.x)); }
    }
}
#define FLD_QMAX 16
struct fld_grid_args { int B, N, P, Q, nx, ny, maps, ysplit; const cd* Sall; const cd* Xt; const cd* Yt; cd* F; };
// QT = Q at compile time (the T registers and the q loops are exact: no predicates, half the registers of the generic form, three
// CTAs per SM); QT = 0: any Q <= FLD_QMAX.
template <int QT>
KH_DEV void fld_grid_body_t(const Cta& c, const fld_grid_args& a) {
    constexpr int QN = QT > 0 ? QT : FLD_QMAX;
    const int b = c.bx, map = c.by / a.ysplit, part = c.by - map * a.ysplit;         // map = iz * 6 + component
    const int P = a.P, Q = QT > 0 ? QT : a.Q, N = a.N, nx = a.nx, ny = a.ny, tid = c.tid, nthr = c.nthr;
    const int y0 = (int)((long long)ny * part / a.ysplit), y1 = (int)((long long)ny * (part + 1) / a.ysplit);
    // shared: [S N][Y Q x (y1 - y0)]
    cd* Ss = (cd*)c.smem;
    cd* Ys = Ss + N;
    const int nyl = y1 - y0;
    const cd* S = a.Sall + ((long long)b * a.maps + map) * N;
    for (int g = tid; g < N; g += nthr) Ss[g] = S[g];
    for (int e = tid; e < Q * nyl; e += nthr) { const int q = e / nyl, i = e - q * nyl; Ys[e] = a.Yt[((long long)b * Q + q) * ny + y0 + i]; }
    c.sync();
    cd* out = a.F + ((long long)b * a.maps + map) * ny * nx;
    for (int ix = tid; ix < nx; ix += nthr) {
        cd T[QN];
#pragma unroll
        for (int q = 0; q < QN; ++q) {
            T[q] = mk(0, 0);
            if (q < Q) {
                const cd* xr = a.Xt + ((long long)b * N + q * P) * nx + ix;
                cd acc = mk(0, 0);
                for (int p = 0; p < P; ++p) cfma(acc, Ss[q * P + p], xr[(long long)p * nx]);
                T[q] = acc;
            }
        }
        cd* o = out + (long long)y0 * nx + ix;
        const cd* yr = Ys;
        int iy = 0;
        for (; iy + 4 <= nyl; iy += 4, yr += 4, o += 4LL * nx) {             // four independent accumulation chains per thread
            cd f0 = mk(0, 0), f1 = mk(0, 0), f2 = mk(0, 0), f3 = mk(0, 0);
#pragma unroll
            for (int q = 0; q < QN; ++q) if (q < Q) {
                const cd* yq = yr + q * nyl;
                cfma(f0, T[q], yq[0]); cfma(f1, T[q], yq[1]); cfma(f2, T[q], yq[2]); cfma(f3, T[q], yq[3]);
            }
            kh_store_stream(o, f0); kh_store_stream(o + nx, f1); kh_store_stream(o + 2LL * nx, f2); kh_store_stream(o + 3LL * nx, f3);
        }
        for (; iy < nyl; ++iy, ++yr, o += nx) {
            cd f0 = mk(0, 0);
#pragma unroll
            for (int q = 0; q < QN; ++q) if (q < Q) cfma(f0, T[q], yr[q * nyl]);
            kh_store_stream(o, f0);
        }
    }
}
KH_DEV void fld_grid_body(const Cta& c, const fld_grid_args& a) { fld_grid_body_t<0>(c, a); }
KH_DEV void fld_grid5_body(const Cta& c, const fld_grid_args& a) { fld_grid_body_t<5>(c, a); }
KH_DEV void fld_grid7_body(const Cta& c, const fld_grid_args& a) { fld_grid_body_t<7>(c, a); }
KH_DEV void fld_grid9_body(const Cta& c, const fld_grid_args& a) { fld_grid_body_t<9>(c, a); }
KH_DEV void fld_grid11_body(const Cta& c, const fld_grid_args& a) { fld_grid_body_t<11>(c, a); }

struct FieldBufs {
    cd *Kx, *Ky; double* k0; cd* c1p; cd* Fm; cd* Finv; cd* y12; cd* m12; cd* Winv; cd* Vinv; cd* Sall; cd* Ph;
    int* zpos; int* zlayer; double* zdist; const cd** ICp; cd* ICs; int* info; int* stack;
    int group;                          // stack positions (or layers) whose inverses go out as ONE batched launch
    cd* zwork; long long zwork_cd;      // work space of the blocked inverse (n beyond shared memory)
    cd* Xt; cd* Yt;                     // grid path: phase tables
};
static void layout_fields(const kh_plan* p, int B, int npts, int nz, Bump& b, FieldBufs& f, int nx = 0, int ny = 0) {
    const size_t N = p->N, n = p->n, n2 = n * n, nL = p->layers.size(), Ls = p->stack.size();
    f.Kx = b.get<cd>(B * N); f.Ky = b.get<cd>(B * N); f.k0 = b.get<double>(B);
    // The inverses of the pipeline (W^-1, V^-1 per layer, F^-1 per stack position) are independent of each other: positions are
    // processed in groups so that a small batch of solves (17 frequencies of a field map) still fills the GPU with one launch
    // per group instead of one launch of B matrices per position.  Up to two waves of one CTA per SM per launch.
    size_t group = B >= 296 ? 1 : 296 / (size_t)B;
    if (group > (Ls > nL ? Ls : nL)) group = (Ls > nL ? Ls : nL);
    f.group = (int)group;
    const size_t Bz = (size_t)B * group;
    f.c1p = b.get<cd>(B * n); f.Fm = b.get<cd>(Bz * n2); f.Finv = b.get<cd>(Bz * n2);
    f.y12 = b.get<cd>(Bz * 2 * n); f.m12 = b.get<cd>(B * Ls * 2 * n);
    f.Winv = b.get<cd>(nL * B * n2); f.Vinv = b.get<cd>(nL * B * n2);      // [B][nL][n][n]
    f.Sall = b.get<cd>((size_t)B * nz * 6 * N);
    if (nx > 0) { f.Ph = nullptr; f.Xt = b.get<cd>((size_t)B * N * nx); f.Yt = b.get<cd>((size_t)B * p->Q * ny); }
    else { f.Ph = b.get<cd>((size_t)B * N * npts); f.Xt = f.Yt = nullptr; }
    f.zpos = b.get<int>(nz); f.zlayer = b.get<int>(nz); f.zdist = b.get<double>(nz);
    f.ICp = b.get<const cd*>(nL); f.ICs = b.get<cd>(nL); f.info = b.get<int>(2 * Bz); f.stack = b.get<int>(Ls);
    f.zwork_cd = (int)n >= KH_ZINV_BLOCKED_MIN ? (long long)Bz * zinv_work_cd((int)n) : 0;
#ifndef KH_HOST_EMU
    if ((int)n >= KH_ZINV_BLOCKED_MIN && (int)n <= ZIL_NMAX && (long long)Bz * zinv_l2_work_cd((int)n) > f.zwork_cd) f.zwork_cd = (long long)Bz * zinv_l2_work_cd((int)n);
#endif
    f.zwork = b.get<cd>((size_t)f.zwork_cd);
}

extern "C" size_t kh_fields_workspace_bytes(const kh_plan* plan, int B, int npts, int nz) {
    if (!plan || B < 1 || npts < 1 || nz < 1) return 0;
    Bump b{nullptr, 0, 0};
    FieldBufs f;
    layout_fields(plan, B, npts, nz, b, f);
    return b.off + 256;
}

static int kh_h2d(void* dst, const void* src, size_t bytes, kh_stream_t st) {
#ifdef KH_HOST_EMU
    (void)st; memcpy(dst, src, bytes); return 0;
#else
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // the host arrays are small temporaries
    return (int)e;
#endif
}

extern "C" size_t kh_fields_grid_workspace_bytes(const kh_plan* plan, int B, int nx, int ny, int nz) {
    if (!plan || B < 1 || nx < 1 || ny < 1 || nz < 1) return 0;
    Bump b{nullptr, 0, 0};
    FieldBufs f;
    layout_fields(plan, B, nx * ny, nz, b, f, nx, ny);
    return b.off + 256;
}

// grid = 0: scattered points (x_dev[p], y_dev[p]), p < npts.  grid = 1: x_dev[nx], y_dev[ny] are the axes of a rectangular grid.
// grid = 2: no inverse transform, F_dev receives the Fourier fields [B][nz][6][N] (x_dev, y_dev unused, npts = 1 for the layout).
static int fields_impl(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                       const kh_outputs* solved, const double* x_dev, const double* y_dev, int npts, int grid, int nx, int ny,
                       const double* z_host, int nz, const double* zpos_host, void* F_dev,
                       void* ws_dev, size_t ws_bytes, void* stream) {
    if (!plan || B < 0 || !wl_dev || !kp_dev || !inc_dev || !solved || (grid != 2 && (!x_dev || !y_dev)) || npts < 1 || !z_host || nz < 1 || !zpos_host || !F_dev || !ws_dev)
        return fail(KH_EINVAL, "kh_fields_batch: bad arguments");
    if (!(solved->prefix_dev && solved->suffix_dev && solved->W_dev && solved->V_dev && solved->L_dev))
        return fail(KH_EINVAL, "kh_fields_batch: needs the KH_WANT_FIELDS outputs of kh_solve_batch");
    if (B == 0) return 0;
    const kh_plan* p = plan;
    if (p->layers[p->stack[0]].kind != KH_LAYER_HALF_INC) return fail(KH_EINVAL, "kh_fields_batch: the stack must start with the incidence half space");
    if (grid == 1 && p->Q > FLD_QMAX) return fail(KH_EINVAL, "kh_fields_grid_batch: more than 16 harmonics along y; use kh_fields_batch");
    if (ws_bytes < (grid == 1 ? kh_fields_grid_workspace_bytes(p, B, nx, ny, nz) : kh_fields_workspace_bytes(p, B, npts, nz)))
        return fail(KH_ENOMEM, "kh_fields_batch: workspace too small");
    kh_stream_t st = (kh_stream_t)stream;
    const int N = p->N, n = p->n, Ls = (int)p->stack.size(), nL = (int)p->layers.size();
    const long long n2 = (long long)n * n;
    Bump bump{(char*)ws_dev, ws_bytes, 0};
    FieldBufs f;
    layout_fields(p, B, npts, nz, bump, f, grid == 1 ? nx : 0, grid == 1 ? ny : 0);
    if (grid == 2) f.Sall = (cd*)F_dev;                      // fld_z writes the coefficients straight into the caller's buffer

    // host: locate every depth (crystal.py:208-232) -> stack position, layer, distance to the right face
    std::vector<int> zpos(nz), zlay(nz);
    std::vector<double> zdist(nz);
    std::vector<char> pos_active(Ls, 0), lay_active(nL, 0);
    for (int k = 0; k < nz; ++k) {
        const double z = z_host[k];
        int idx = 0;
        while (idx < Ls + 1 && zpos_host[idx] < z) ++idx;        // numpy.searchsorted(side='left')
        idx -= 1;
        if (idx < 0) idx = 0;
        if (idx > Ls - 1) idx = Ls - 1;
        const double zr = (z <= 0.0) ? z : z - zpos_host[idx];
        zpos[k] = idx; zlay[k] = p->stack[idx];
        zdist[k] = p->layers[zlay[k]].depth - zr;
        pos_active[idx] = 1; lay_active[zlay[k]] = 1;
    }
    std::vector<const cd*> icp(nL, nullptr);
    std::vector<cd> ics(nL, mk(1, 0));
    for (int i = 0; i < nL; ++i) {
        const kh_layer_desc& L = p->layers[i];
        if (L.kind == KH_LAYER_PIXMAP) icp[i] = (const cd*)L.IC_dev;
        else if (L.kind == KH_LAYER_EXTENDED) ics[i] = mk(1.0, 0.0);            // ExtendedLayer.IC = 1.0 (extension.py:112)
        else ics[i] = crecip(mk(L.eps_re, L.eps_im));
    }
    KH_TRY(kh_h2d(f.zpos, zpos.data(), nz * sizeof(int), st));
    KH_TRY(kh_h2d(f.zlayer, zlay.data(), nz * sizeof(int), st));
    KH_TRY(kh_h2d(f.zdist, zdist.data(), nz * sizeof(double), st));
    KH_TRY(kh_h2d(f.ICp, icp.data(), nL * sizeof(cd*), st));
    KH_TRY(kh_h2d(f.ICs, ics.data(), nL * sizeof(cd), st));

    {   kvec_args a{B, N, wl_dev, (const cd*)kp_dev, p->g_dev, f.Kx, f.Ky, f.k0};
        KH_TRY((kh_launch<kvec_args, kvec_body>(dim3(B), 128, 0, st, a))); }
    {   const kh_layer_desc& R = p->layers[p->stack[0]];
        fld_c1p_args a{B, N, mk(R.eps_re, R.eps_im), f.Kx, f.Ky, (const cd*)inc_dev, f.c1p};
        KH_TRY((kh_launch<fld_c1p_args, fld_c1p_body>(dim3(B), 128, 0, st, a))); }

    cd* Wd = (cd*)solved->W_dev; cd* Vd = (cd*)solved->V_dev;
    KH_TRY(kh_h2d(f.stack, p->stack.data(), Ls * sizeof(int), st));
    // W^-1, V^-1 of the layers that hold a depth: runs of consecutive active layers, at most f.group per launch; batch index
    // b * run + j <-> (solve b, layer l0 + j) on the [B][nL][n][n] stacks
    for (int l0 = 0; l0 < nL;) {
        if (!lay_active[l0]) { ++l0; continue; }
        int run = 1;
        while (l0 + run < nL && lay_active[l0 + run] && run < f.group) ++run;
        KH_TRY(zinv_launch(st, B * run, n, mref(Wd + (long long)l0 * n2, (long long)nL * n2, n, run, n2),
                           mref(f.Winv + (long long)l0 * n2, (long long)nL * n2, n, run, n2), f.info, f.zwork, f.zwork_cd));
        KH_TRY(zinv_launch(st, B * run, n, mref(Vd + (long long)l0 * n2, (long long)nL * n2, n, run, n2),
                           mref(f.Vinv + (long long)l0 * n2, (long long)nL * n2, n, run, n2), f.info, f.zwork, f.zwork_cd));
        l0 += run;
    }
    cd* pre = (cd*)solved->prefix_dev; cd* suf = (cd*)solved->suffix_dev;
    const long long sstride = (long long)Ls * 4 * n2;
    for (int i0 = 0; i0 < Ls;) {
        if (!pos_active[i0]) { ++i0; continue; }
        int run = 1;
        while (i0 + run < Ls && pos_active[i0 + run] && run < f.group) ++run;
        MatRef Sl22 = mref(pre + ((long long)i0 * 4 + 3) * n2, sstride, n, run, 4 * n2), Sl21 = mref(pre + ((long long)i0 * 4 + 2) * n2, sstride, n, run, 4 * n2);
        MatRef Sr11 = mref(suf + ((long long)i0 * 4 + 0) * n2, sstride, n, run, 4 * n2);
        MatRef Fm = mref(f.Fm, n2, n), Fi = mref(f.Finv, n2, n);
        KH_TRY(gemm(st, B * run, n, Sl22, Sr11, Fm, -1.0, nullptr, 0.0, 1.0));
        KH_TRY(zinv_launch(st, B * run, n, Fm, Fi, f.info, f.zwork, f.zwork_cd));
        {   fld_amp_args a{B, N, run, Fi, Sl21, Sr11, f.c1p, f.Kx, f.Ky, f.y12};
            KH_TRY((kh_launch<fld_amp_args, fld_amp_body>(dim3(B * run), 256, (size_t)3 * n * sizeof(cd), st, a))); }
        {   fld_modes_args a{B, n, run, i0, nL, f.stack, f.Winv, f.Vinv, f.y12, f.m12, (long long)Ls * 2 * n};
            KH_TRY((kh_launch<fld_modes_args, fld_modes_body>(dim3(B * run), 256, (size_t)2 * n * sizeof(cd), st, a))); }
        i0 += run;
    }
    {   fld_z_args a{B, N, nz, f.zpos, f.zlayer, f.zdist, f.m12, (long long)Ls * 2 * n, 2LL * n,
                     Wd, Vd, (const cd*)solved->L_dev, (long long)nL * n2, nL, f.ICp, f.ICs, f.Kx, f.Ky, f.k0, f.Sall};
        KH_TRY((kh_launch<fld_z_args, fld_z_body>(dim3(B, nz), 256, (size_t)5 * n * sizeof(cd), st, a))); }
    if (grid == 2) return 0;
    if (grid) {
        {   fld_gtab_args a{B, N, p->P, p->Q, nx, ny, (const cd*)kp_dev, p->g_dev, x_dev, y_dev, f.Xt, f.Yt};
            KH_TRY((kh_launch<fld_gtab_args, fld_gtab_body>(dim3(B, N + p->Q), 256, 0, st, a, "fld_grid"))); }
        // enough CTAs for a few waves: split the y range when the batch of maps is small
        int ysplit = 1;
        while ((long long)B * nz * 6 * ysplit < 4 * 148 && ysplit * 2 <= ny && ysplit < 16) ysplit *= 2;
        // ... and until the y phase table of one CTA fits in shared memory (tall grids: ny = 2048 with Q = 9 needs 295 KB unsplit)
        auto table_bytes = [&](int ys) { return ((size_t)N + (size_t)p->Q * ((ny + ys - 1) / ys + 1)) * sizeof(cd); };
        while (table_bytes(ysplit) > (size_t)KH_SMEM_MAX && ysplit < ny) ysplit *= 2;
        fld_grid_args a{B, N, p->P, p->Q, nx, ny, 6 * nz, ysplit, f.Sall, f.Xt, f.Yt, (cd*)F_dev};
        const size_t sm = table_bytes(ysplit);
        if (sm > (size_t)KH_SMEM_MAX) return fail(KH_EINVAL, "kh_fields_grid_batch: basis too large for the phase table in shared memory");
        const double work = 8.0 * ((double)N * nx + (double)p->Q * nx * ny) * 6.0 * nz * B;
        const dim3 gg(B, 6 * nz * ysplit);
        const int thr = nx >= 256 ? 256 : (nx >= 128 ? 128 : 64);
        switch (p->Q) {          // the usual odd bases with the q loops exact; anything else through the generic form
            case 5: return kh_launch<fld_grid_args, fld_grid5_body, 256, 3>(gg, thr, sm, st, a, "fld_grid", work);
            case 7: return kh_launch<fld_grid_args, fld_grid7_body, 256, 3>(gg, thr, sm, st, a, "fld_grid", work);
            case 9: return kh_launch<fld_grid_args, fld_grid9_body, 256, 3>(gg, thr, sm, st, a, "fld_grid", work);
            case 11: return kh_launch<fld_grid_args, fld_grid11_body, 256, 2>(gg, thr, sm, st, a, "fld_grid", work);
            default: return kh_launch<fld_grid_args, fld_grid_body>(gg, thr, sm, st, a, "fld_grid", work);
        }
    }
    {   fld_phase_args a{B, N, npts, (const cd*)kp_dev, p->g_dev, x_dev, y_dev, f.Ph};
        KH_TRY((kh_launch<fld_phase_args, fld_phase_body>(dim3(B, N), 256, 0, st, a))); }
    {   zgemm_args g = zgemm_make(6 * nz, npts, N, mref(f.Sall, (long long)nz * 6 * N, N), mref(f.Ph, (long long)N * npts, npts),
                                  mref(F_dev, (long long)nz * 6 * npts, npts));
        KH_TRY(zgemm_launch(st, B, g)); }
    return 0;
}

extern "C" int kh_fields_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                               const kh_outputs* solved, const double* x_dev, const double* y_dev, int npts,
                               const double* z_host, int nz, const double* zpos_host, void* F_dev,
                               void* ws_dev, size_t ws_bytes, void* stream) {
    return fields_impl(plan, B, wl_dev, kp_dev, inc_dev, solved, x_dev, y_dev, npts, 0, 0, 0, z_host, nz, zpos_host, F_dev, ws_dev, ws_bytes, stream);
}
extern "C" int kh_fields_grid_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                                    const kh_outputs* solved, const double* xs_dev, int nx, const double* ys_dev, int ny,
                                    const double* z_host, int nz, const double* zpos_host, void* F_dev,
                                    void* ws_dev, size_t ws_bytes, void* stream) {
    if (nx < 1 || ny < 1) return fail(KH_EINVAL, "kh_fields_grid_batch: bad grid");
    return fields_impl(plan, B, wl_dev, kp_dev, inc_dev, solved, xs_dev, ys_dev, nx * ny, 1, nx, ny, z_host, nz, zpos_host, F_dev, ws_dev, ws_bytes, stream);
}
extern "C" int kh_fields_fourier_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                                       const kh_outputs* solved, const double* z_host, int nz, const double* zpos_host, void* S_dev,
                                       void* ws_dev, size_t ws_bytes, void* stream) {
    return fields_impl(plan, B, wl_dev, kp_dev, inc_dev, solved, nullptr, nullptr, 1, 2, 0, 0, z_host, nz, zpos_host, S_dev, ws_dev, ws_bytes, stream);
}

// ---- fourier.idft (khepri/fourier.py:136-142) as a standalone operator: out[m][p] = sum_g s[m][g] exp(i (kx_g x_p + ky_g y_p)),
// kx, ky complex per harmonic (callers pass k0 * Kx), points scattered.  Phase matrix, then ONE DMMA GEMM [M x N] . [N x npts].
struct idft_phase_args { int N, npts, gy; const cd* kx; const cd* ky; const double* x; const double* y; cd* Ph; };
KH_DEV void idft_phase_body(const Cta& c, const idft_phase_args& a) {
    const int g = c.bx;
    const cd kx = a.kx[g], ky = a.ky[g];
    cd* o = a.Ph + (long long)g * a.npts;
    for (int p = c.by * c.nthr + c.tid; p < a.npts; p += c.nthr * a.gy) {
        cd arg = a.x[p] * kx + a.y[p] * ky;
        o[p] = cexp_(mk(-arg.y, arg.x));
    }
}
extern "C" size_t kh_idft_work_bytes(int N, int npts) { return (size_t)N * npts * sizeof(cd) + 512; }
extern "C" int kh_idft_batch(int M, int N, int npts, const void* kx_dev, const void* ky_dev, const double* x_dev, const double* y_dev,
                             const void* s_dev, void* out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
    if (M < 0 || N < 1 || npts < 1 || !kx_dev || !ky_dev || !x_dev || !y_dev || !s_dev || !out_dev || !ws_dev)
        return fail(KH_EINVAL, "kh_idft_batch: bad arguments");
    if (ws_bytes < kh_idft_work_bytes(N, npts)) return fail(KH_ENOMEM, "kh_idft_batch: workspace too small");
    if (M == 0) return 0;
    kh_stream_t st = (kh_stream_t)stream;
    Bump bump{(char*)ws_dev, ws_bytes, 0};
    cd* Ph = bump.get<cd>((size_t)N * npts);
    int gy = (npts + 4095) / 4096; if (gy > 64) gy = 64;
    {   idft_phase_args a{N, npts, gy, (const cd*)kx_dev, (const cd*)ky_dev, x_dev, y_dev, Ph};
        KH_TRY((kh_launch<idft_phase_args, idft_phase_body>(dim3(N, gy), 256, 0, st, a, "idft_phase"))); }
    zgemm_args g = zgemm_make(M, npts, N, mref((cd*)s_dev, 0, N), mref(Ph, 0, npts), mref(out_dev, 0, npts));
    KH_TRY(zgemm_launch(st, 1, g));
    return 0;
}

// ---- Brillouin-zone-integration source (khepri/beams.py:164-191, amplitudes_from_fields): Fourier amplitudes of a
// real-space beam for every k-point of the BZ grid,
//     amp[b][g][c] = scale * sum_p F[p][c] exp(-i ((kp_x[b] + g_x) x_p + (kp_y[b] + g_y) y_p)),   c = (Ex, Ey, Hx, Hy)
// (the reference divides the samples by the Bloch phase of k and calls slow_dft per supercell tile; the sum over tiles is
// one sum over all samples).  Phase matrix by fld_phase (sign -1), then ONE DMMA GEMM  [N x npts] . [npts x 4]  per k-point
// with the sample matrix shared by the whole batch.
extern "C" size_t kh_beam_amplitudes_work_bytes(int B, int N, int npts) {
    return (size_t)B * N * npts * sizeof(cd) + 512;
}
extern "C" int kh_beam_amplitudes(int B, int N, int npts, const void* kp_dev, const double* g_dev, const double* x_dev, const double* y_dev,
                                  const void* fields_dev, double scale, void* amp_dev, void* ws_dev, size_t ws_bytes, void* stream) {
    if (B < 0 || N < 1 || npts < 1 || !kp_dev || !g_dev || !x_dev || !y_dev || !fields_dev || !amp_dev || !ws_dev)
        return fail(KH_EINVAL, "kh_beam_amplitudes: bad arguments");
    if (ws_bytes < kh_beam_amplitudes_work_bytes(B, N, npts)) return fail(KH_ENOMEM, "kh_beam_amplitudes: workspace too small");
    if (B == 0) return 0;
    kh_stream_t st = (kh_stream_t)stream;
    Bump bump{(char*)ws_dev, ws_bytes, 0};
    cd* Ph = bump.get<cd>((size_t)B * N * npts);
    {   fld_phase_args a{B, N, npts, (const cd*)kp_dev, g_dev, x_dev, y_dev, Ph, -1.0};
        KH_TRY((kh_launch<fld_phase_args, fld_phase_body>(dim3(B, N), 256, 0, st, a, "beam_phase"))); }
    zgemm_args g = zgemm_make(N, 4, npts, mref(Ph, (long long)N * npts, npts), mref(fields_dev, 0, 4), mref(amp_dev, (long long)N * 4, 4), scale);
    KH_TRY(zgemm_launch(st, B, g));
    return 0;
}
