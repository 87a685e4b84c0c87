// Extended (twisted-bilayer) RCWA support, retained eigenspaces and the reverse chain.
// Included by kh_api.cu after the chunk layout helpers.
//
// khepri/extension.py:82-112: an ExtendedLayer solves its base layer at the N_b k-points
// kp + g_shift (g_shift = the *other* lattice's reciprocal vectors) and scatters the N_b small
// S-matrices (2N_b x 2N_b) into the moire basis (2N_b^2 x 2N_b^2).  Here the N_b * Bc sub-solves
// are just a bigger batch for the same kernels, and the scatter is one gather-style kernel
// (every output element is written exactly once, so no memset is needed).
#pragma once

struct ext_kp_args { int Bc, Nb; const double* wl; const cd* kp; const double* gshift; double* wl_sub; cd* kp_sub; };
KH_DEV void ext_kp_body(const Cta& c, const ext_kp_args& a) {
    const int b = c.bx;
    for (int s = c.tid; s < a.Nb; s += c.nthr) {
        long long o = (long long)b * a.Nb + s;
        a.wl_sub[o] = a.wl[b];
        a.kp_sub[2 * o] = mk(a.kp[2 * b].x + a.gshift[s], a.kp[2 * b].y);
        a.kp_sub[2 * o + 1] = mk(a.kp[2 * b + 1].x + a.gshift[a.Nb + s], a.kp[2 * b + 1].y);
    }
}

// joint[b][blk][row][col] from sub-solve S-matrices.  src_kind: 0 = BD tables [Bc*Nb][16 Nb],
// 1 = symmetric dense [Bc*Nb][2][nb][nb].  mode as extension.py:35-38.
struct ext_scatter_args { int Bc, Nb, mode, src_kind; const cd* src; cd* dst; };
KH_DEV void ext_scatter_body(const Cta& c, const ext_scatter_args& a) {
    const int Nb = a.Nb, nb = 2 * Nb, NN = Nb * Nb, n = 2 * NN, blk = c.by, b = c.bx;
    cd* d = a.dst + ((long long)b * 4 + blk) * n * n;
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int row = e / n, col = e - row * n;
        int ha = row >= NN, hb = col >= NN;
        int jr = row - ha * NN, jc = col - hb * NN;
        int sr, r1, sc, r2;
        if (a.mode == 0) { sr = jr / Nb; r1 = jr % Nb; sc = jc / Nb; r2 = jc % Nb; }
        else { r1 = jr / Nb; sr = jr % Nb; r2 = jc / Nb; sc = jc % Nb; }
        cd v = mk(0, 0);
        if (sr == sc) {
            long long sub = (long long)b * Nb + sr;
            if (a.src_kind == 0) {
                if (r1 == r2) v = a.src[sub * 16 * Nb + (long long)blk * 4 * Nb + (ha * 2 + hb) * Nb + r1];
            } else {
                const int map[4] = {0, 1, 1, 0};
                v = a.src[(sub * 2 + map[blk]) * nb * nb + (long long)(ha * Nb + r1) * nb + hb * Nb + r2];
            }
        }
        d[e] = v;
    }
}

// Retained eigenspace of an extended layer (extension.py:100-110): W, V and lambda of the N_b shifted base solves scattered into
// the joint basis with the same index map as the S-matrix blocks (_joint_subspace, extension.py:10-38); zero between shifts.
// src_kind 0: uniform base (W = I, V as BD table [Be][4][Nb], lambda [Be][Nb]); 1: pixmap base (dense [Be][nb][nb], lambda [Be][nb]).
struct ext_keep_args { int Bc, Nb, mode, src_kind, nL, li; const cd* Wsub; const cd* Vsub; const cd* Lsub; cd* Wout; cd* Vout; cd* Lout; };
KH_DEV void ext_keep_body(const Cta& c, const ext_keep_args& a) {
    const int Nb = a.Nb, nb = 2 * Nb, NN = Nb * Nb, n = 2 * NN, b = c.bx, which = c.by;      // which: 0 -> W, 1 -> V
    const long long o = (long long)b * a.nL + a.li;
    cd* d = (which ? a.Vout : a.Wout) + o * n * n;
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int row = e / n, col = e - row * n;
        int ha = row >= NN, hb = col >= NN;
        int jr = row - ha * NN, jc = col - hb * NN;
        int sr, r1, sc, r2;
        if (a.mode == 0) { sr = jr / Nb; r1 = jr % Nb; sc = jc / Nb; r2 = jc % Nb; }
        else { r1 = jr / Nb; sr = jr % Nb; r2 = jc / Nb; sc = jc % Nb; }
        cd v = mk(0, 0);
        if (sr == sc) {
            const long long sub = (long long)b * Nb + sr;
            if (a.src_kind == 1) v = (which ? a.Vsub : a.Wsub)[sub * nb * nb + (long long)(ha * Nb + r1) * nb + hb * Nb + r2];
            else if (r1 == r2) v = which ? a.Vsub[sub * 4 * Nb + (ha * 2 + hb) * Nb + r1] : mk(ha == hb ? 1.0 : 0.0, 0.0);
        }
        d[e] = v;
    }
    if (which == 0) {
        cd* Lo = a.Lout + o * n;
        for (int i = c.tid; i < n; i += c.nthr) {
            const int ha = i >= NN, jr = i - ha * NN;
            const int sr = a.mode == 0 ? jr / Nb : jr % Nb, r1 = a.mode == 0 ? jr % Nb : jr / Nb;
            const long long sub = (long long)b * Nb + sr;
            Lo[i] = a.src_kind == 1 ? a.Lsub[sub * nb + ha * Nb + r1] : a.Lsub[sub * Nb + r1];
        }
    }
}

static int solve_extended(kh_stream_t st, const kh_plan* p, int Bc, int li, const double* wl, const cd* kp, ChunkBufs& cb, int* info_out,
                          const kh_outputs* keep = nullptr, int b0 = 0) {
    const kh_layer_desc& L = p->layers[li];
    const kh_layer_desc& base = p->layers[L.ext_base];
    const int Nb = p->Nb, Be = Bc * Nb;
    // extension.py:66-80: mode 1 -> base lives on the lhs lattice, shifts are the rhs g-vectors; mode 0 the other way round
    const double* gbase = L.ext_mode == 1 ? p->glhs_dev : p->grhs_dev;
    const double* gshift = L.ext_mode == 1 ? p->grhs_dev : p->glhs_dev;
    double* wl_sub = cb.ewl;
    cd* kp_sub = cb.ekp;
    {   ext_kp_args a{Bc, Nb, wl, kp, gshift, wl_sub, kp_sub};
        KH_TRY((kh_launch<ext_kp_args, ext_kp_body>(dim3(Bc), 64, 0, st, a))); }
    {   kvec_args a{Be, Nb, wl_sub, kp_sub, gbase, cb.eKx, cb.eKy, cb.ek0};
        KH_TRY((kh_launch<kvec_args, kvec_body>(dim3(Be), 64, 0, st, a))); }
    int src_kind;
    const cd* src;
    if (base.kind == KH_LAYER_UNIFORM) {
        bd_layer_args a{Be, Nb, KH_LAYER_UNIFORM, mk(base.eps_re, base.eps_im), base.depth, cb.eKx, cb.eKy, cb.ek0, cb.ebd, keep ? cb.eVbd : nullptr, keep ? cb.eLbd : nullptr};
        KH_TRY((kh_launch<bd_layer_args, bd_layer_body>(dim3(Be), 64, 0, st, a)));
        src_kind = 0; src = cb.ebd;
    } else if (base.kind == KH_LAYER_PIXMAP) {
        LayerVec v = cb.evec; v.info_acc = info_out ? info_out : cb.info; v.info_div = Nb;     // sub-solve b * Nb + s reports into solve b
        KH_TRY(solve_patterned(st, Be, Nb, (const cd*)base.C_dev, (const cd*)base.IC_dev, base.depth, cb.eKx, cb.eKy, cb.ek0,
                               cb.epool, v, cb.eS, keep ? cb.eW : nullptr, keep ? cb.eV : nullptr, keep ? cb.eL : nullptr, (long long)(2 * Nb) * (2 * Nb), 2LL * Nb));
        src_kind = 1; src = cb.eS;
    } else return fail(KH_EINVAL, "extended layer: base must be uniform or pixmap");
    ext_scatter_args a{Bc, Nb, L.ext_mode, src_kind, src, cb.layerS[li]};
    KH_TRY((kh_launch<ext_scatter_args, ext_scatter_body>(dim3(Bc, 4), 256, 0, st, a)));
    if (keep) {
        const int nL = (int)p->layers.size();
        const long long n2 = (long long)p->n * p->n;
        ext_keep_args k{Bc, Nb, L.ext_mode, src_kind, nL, li, src_kind ? cb.eW : nullptr, src_kind ? cb.eV : cb.eVbd, src_kind ? cb.eL : cb.eLbd,
                        (cd*)keep->W_dev + (long long)b0 * nL * n2, (cd*)keep->V_dev + (long long)b0 * nL * n2, (cd*)keep->L_dev + (long long)b0 * nL * p->n};
        KH_TRY((kh_launch<ext_keep_args, ext_keep_body>(dim3(Bc, 2), 256, 0, st, k)));
    }
    return 0;
}

// retained eigenspace of BD layers, expanded to the dense per-layer format: W = I, V (2x2 blocks of
// diagonals), lambda duplicated for both halves
struct bd_keep_args { int Bc, N, nL, li; const cd* V; const cd* lam; cd* Wout; cd* Vout; cd* Lout; };
KH_DEV void bd_keep_body(const Cta& c, const bd_keep_args& a) {
    const int N = a.N, n = 2 * N, b = c.bx;
    const long long o = ((long long)b * a.nL + a.li);
    cd* W = a.Wout + o * n * n; cd* V = a.Vout + o * n * n; cd* Lo = a.Lout + o * n;
    const cd* v = a.V + (long long)b * 4 * N;
    for (int e = c.tid; e < n * n; e += c.nthr) {
        int i = e / n, j = e - i * n;
        int gi = i < N ? i : i - N, gj = j < N ? j : j - N;
        W[e] = mk(i == j ? 1.0 : 0.0, 0.0);
        V[e] = (gi == gj) ? v[((i >= N) * 2 + (j >= N)) * N + gi] : mk(0, 0);
    }
    for (int i = c.tid; i < n; i += c.nthr) Lo[i] = a.lam[(long long)b * N + (i < N ? i : i - N)];
}

static int keep_eigenspace_bd(kh_stream_t st, const kh_plan* p, int Bc, ChunkBufs& cb, const kh_outputs* out, int b0) {
    const int nL = (int)p->layers.size();
    const long long n2 = (long long)p->n * p->n;
    for (int i = 0; i < nL; ++i) {
        if (!layer_is_bd(p, i) || !cb.layerV[i]) continue;
        bool used = false;
        for (int s : p->stack) used |= (s == i);
        if (!used) continue;
        bd_keep_args a{Bc, p->N, nL, i, cb.layerV[i], cb.layerL[i], (cd*)out->W_dev + (long long)b0 * nL * n2,
                       (cd*)out->V_dev + (long long)b0 * nL * n2, (cd*)out->L_dev + (long long)b0 * nL * p->n};
        KH_TRY((kh_launch<bd_keep_args, bd_keep_body>(dim3(Bc), 256, 0, st, a)));
    }
    return 0;
}

// reverse partial products (layer.py:49-59): suffix[Ls-1] = identity, suffix[i] = S_{i+1} (*) suffix[i+1]
static int reverse_chain(kh_stream_t st, const kh_plan* p, int Bc, const std::vector<SRef>& S, ChunkBufs& cb, const kh_outputs* out, int b0, int* info_acc) {
    const int N = p->N, n = p->n, Ls = (int)p->stack.size();
    const long long n2 = (long long)n * n;
    {   bd_identity_args a{Bc, N, cb.accB[0]};
        KH_TRY((kh_launch<bd_identity_args, bd_identity_body>(dim3(Bc), 128, 0, st, a))); }
    SRef acc = sref_bd(cb.accB[0]);
    int pb = 1, pd = 0;
    for (int i = Ls - 1; i >= 0; --i) {
        KH_TRY(materialise(st, Bc, N, acc, (cd*)out->suffix_dev + ((long long)b0 * Ls + i) * 4 * n2, (long long)Ls * 4 * n2, cb.expA));
        if (i == 0) break;
        const SRef& L = S[p->stack[i]];
        SRef r2;
        KH_TRY(star_any(st, Bc, N, L, acc, cb.accR[pd], cb.accB[pb], cb.pool, info_acc, r2, -1, false, 1));
        if (r2.bd) pb ^= 1; else pd ^= 1;
        acc = r2;
    }
    return 0;
}
