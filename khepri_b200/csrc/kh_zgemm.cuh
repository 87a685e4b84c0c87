// Batched complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64).
//
//   Cout[b] = rs[i] * ( alpha * op(A[b]) * B[b] + beta * Cin[b] + diag * I ) (*|/) cs[j]
//
// tcgen05 has no FP64 kind, so warp-level DMMA is the tensor path for complex128 on sm_100a.
// A complex product is four real DMMAs per 8x8x4 step (Cr += Ar Br - Ai Bi ; Ci += Ar Bi + Ai Br).
// CTA tile 64x64 or 56x56, K chunks of 16 staged with cp.async (zero-filled at the ragged edges, so
// matrices need no padding in HBM) in a 3-stage pipeline with one barrier per chunk.  64x64: eight warps,
// warp w owns rows 8w..8w+7 of the tile and all eight 8-column MMA tiles.  56x56: the 49 (strip, tile)
// units are dealt evenly over 8 warps (zgemm_body_u).  Shared-memory strides (20 / 66 complex) make the
// A- and B-fragment loads bank-conflict free.
#pragma once
#include "kh_common.cuh"

struct zgemm_args {
    int M, N, K;
    int transA;                 // 1: A is stored K x M (row-major) and used transposed
    MatRef A, B, Cin, Cout;     // Cin.p may be null
    double alpha, beta, diag;
    const cd* rowscale;         // optional, length M per batch group
    long long rs_stride; int rs_group;
    const cd* colscale;         // optional, length N per batch group
    long long cs_stride; int cs_group;
    int cs_divide;              // 1: divide by colscale instead of multiplying
};

#define ZG_BK 16
#ifndef KH_ZG_GROUP
#define KH_ZG_GROUP 3           /* units whose fragments are in flight at once in the unit-balanced kernel */
#endif
#define ZG_LDA 20
#define ZG_EMU_TILE 64

KH_DEV cd zgemm_epilogue(const zgemm_args& a, const cd* cin, const cd* rs, const cd* cs, int row, int col, cd acc) {
    cd v = a.alpha * acc;
    if (cin) v = v + a.beta * cin[(long long)row * a.Cin.ld + col];
    if (a.diag != 0.0 && row == col) v.x += a.diag;
    if (rs) v = rs[row] * v;
    if (cs) v = a.cs_divide ? v / cs[col] : v * cs[col];
    return v;
}

#ifndef KH_HOST_EMU
__device__ __forceinline__ void kh_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void kh_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void kh_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void kh_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

// Host-emulation stand-in for the DMMA kernels below (tests/hostemu): the same tiling of the output, plain loops.
template <int NW, int NT>
KH_DEV void zgemm_body_t(const Cta& c, const zgemm_args& a) {
    constexpr int BM = 8 * NW, BN = 8 * NT;
    const int tiles_n = (a.N + BN - 1) / BN, tiles = ((a.M + BM - 1) / BM) * tiles_n;
    const int b = c.bx / tiles, tile = c.bx - b * tiles;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const cd* A = mat_ptr(a.A, b);
    const cd* B = mat_ptr(a.B, b);
    const cd* Cin = mat_ptr(a.Cin, b);
    cd* Cout = mat_ptr(a.Cout, b);
    const cd* rs = a.rowscale ? a.rowscale + (long long)(b / a.rs_group) * a.rs_stride : (const cd*)0;
    const cd* cs = a.colscale ? a.colscale + (long long)(b / a.cs_group) * a.cs_stride : (const cd*)0;
    for (int i = m0; i < m0 + BM && i < a.M; ++i)
        for (int j = n0; j < n0 + BN && j < a.N; ++j) {
            cd acc = mk(0, 0);
            for (int k = 0; k < a.K; ++k) {
                cd av = a.transA ? A[(long long)k * a.A.ld + i] : A[(long long)i * a.A.ld + k];
                cfma(acc, av, B[(long long)k * a.B.ld + j]);
            }
            Cout[(long long)i * a.Cout.ld + j] = zgemm_epilogue(a, Cin, rs, cs, i, j, acc);
        }
}

#ifndef KH_HOST_EMU
// One full K chunk (4 DMMA k-steps) on NA of the warp's NT column tiles, branch free.
template <int NT, int NA, int LDB>
__device__ __forceinline__ void zgemm_mma_chunk(double (&cr)[NT][2], double (&ci)[NT][2], const cd* as, const cd* bs) {
#pragma unroll
    for (int kk = 0; kk < ZG_BK / 4; ++kk) {
        const cd av = as[kk * 4];
        const double nai = -av.y;
        constexpr int G = NA > 8 ? (NA + 1) / 2 : NA;      // B fragments held at once (register budget)
#pragma unroll
        for (int g0 = 0; g0 < NA; g0 += G) {
            cd bv[G];
#pragma unroll
            for (int t = 0; t < G; ++t) if (g0 + t < NA) bv[t] = bs[kk * 4 * LDB + (g0 + t) * 8];
#pragma unroll
            for (int t = 0; t < G; ++t) if (g0 + t < NA) { kh_dmma(cr[g0 + t][0], cr[g0 + t][1], av.x, bv[t].x); kh_dmma(ci[g0 + t][0], ci[g0 + t][1], av.x, bv[t].y); }
#pragma unroll
            for (int t = 0; t < G; ++t) if (g0 + t < NA) { kh_dmma(cr[g0 + t][0], cr[g0 + t][1], nai, bv[t].y); kh_dmma(ci[g0 + t][0], ci[g0 + t][1], av.y, bv[t].x); }
        }
    }
}
#endif

// Pipelined variant (the one the launcher uses): NW == NT, so every staging pass is 4 uniform strides per thread
// (no per-element index arithmetic), ST cp.async stages with ONE barrier per K chunk, and a branch-free inner
// loop on interior tiles (the two DMMAs that feed one accumulator are issued NT instructions apart).
template <int NW, int NT, int ST>
KH_DEV void zgemm_body_p(const Cta& c, const zgemm_args& a) {
#ifdef KH_HOST_EMU
    zgemm_body_t<NW, NT>(c, a);
#else
    static_assert(NW == NT, "staging strides assume a square CTA tile");
    constexpr int BM = 8 * NW, BN = 8 * NT, LDB = BN + 2, STAGE = BM * ZG_LDA + ZG_BK * LDB;
    const int tiles_n = (a.N + BN - 1) / BN, tiles = ((a.M + BM - 1) / BM) * tiles_n;
    const int b = c.bx / tiles, tile = c.bx - b * tiles;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const cd* A = mat_ptr(a.A, b);
    const cd* B = mat_ptr(a.B, b);
    cd* sm = (cd*)KH_SMEM(c);                           // [ST] x { A tile [BM][LDA], B tile [BK][LDB] }
    const int tid = c.tid, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int nk = (a.K + ZG_BK - 1) / ZG_BK;
    const int nt = min(NT, (a.N - n0 + 7) >> 3);
    const bool warp_active = (m0 + warp * 8) < a.M;

    // staging roles (4 passes each).  A: thread -> (row am + 2NW*i, k ak), or transposed (row am, k ak + 4i).  B: (k bk + 4i, col bn).
    const int am = a.transA ? tid % BM : tid >> 4, ak = a.transA ? tid / BM : tid & 15;
    const int bk = tid / BN, bn = tid - bk * BN;
    const cd* asrc = a.transA ? A + (long long)ak * a.A.ld + m0 + am : A + (long long)(m0 + am) * a.A.ld + ak;
    const cd* bsrc = B + (long long)bk * a.B.ld + n0 + bn;
    const int adst = am * ZG_LDA + ak, bdst = BM * ZG_LDA + bk * LDB + bn;
    const bool bn_ok = (n0 + bn) < a.N;
    const long long a_pass = a.transA ? 4LL * a.A.ld : 2LL * NW * a.A.ld, b_pass = 4LL * a.B.ld;
    const long long a_chunk = a.transA ? (long long)ZG_BK * a.A.ld : ZG_BK, b_chunk = (long long)ZG_BK * a.B.ld;

    auto stage = [&](int buf, int kc) {
        cd* s = sm + buf * STAGE;
        const int k0 = kc * ZG_BK;
        if (!a.transA) {
            const bool kok = (k0 + ak) < a.K;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = kok && (m0 + am + 2 * NW * i) < a.M;
                kh_cp_async16(s + adst + 2 * NW * i * ZG_LDA, ok ? asrc + i * a_pass : A, ok);
            }
        } else {
            const bool mok = (m0 + am) < a.M;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = mok && (k0 + ak + 4 * i) < a.K;
                kh_cp_async16(s + adst + 4 * i, ok ? asrc + i * a_pass : A, ok);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = bn_ok && (k0 + bk + 4 * i) < a.K;
            kh_cp_async16(s + bdst + 4 * i * LDB, ok ? bsrc + i * b_pass : B, ok);
        }
        asrc += a_chunk; bsrc += b_chunk;
    };

    double cr[NT][2], ci[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) { cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < ST - 1; ++s) { if (s < nk) stage(s, s); kh_cp_async_commit(); }
    int buf = 0, nbuf = ST - 1;
    for (int kc = 0; kc < nk; ++kc) {
        kh_cp_async_wait<ST - 2>();
        __syncthreads();
        if (kc + ST - 1 < nk) stage(nbuf, kc + ST - 1);
        kh_cp_async_commit();
        if (warp_active) {
            const cd* as = sm + buf * STAGE + (warp * 8 + lr) * ZG_LDA + lk;
            const cd* bs = sm + buf * STAGE + BM * ZG_LDA + lk * LDB + lr;
            const int ks = min(ZG_BK / 4, (a.K - kc * ZG_BK + 3) >> 2);
            if (ks == ZG_BK / 4 && nt == NT) zgemm_mma_chunk<NT, NT, LDB>(cr, ci, as, bs);
            else if (ks == ZG_BK / 4 && nt == NT - 1) zgemm_mma_chunk<NT, NT - 1, LDB>(cr, ci, as, bs);
            else {
                for (int kk = 0; kk < ks; ++kk) {
                    const cd av = as[kk * 4];
                    const double nai = -av.y;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        if (t < nt) {
                            const cd bv = bs[kk * 4 * LDB + t * 8];
                            kh_dmma(cr[t][0], cr[t][1], av.x, bv.x);
                            kh_dmma(ci[t][0], ci[t][1], av.x, bv.y);
                            kh_dmma(cr[t][0], cr[t][1], nai, bv.y);
                            kh_dmma(ci[t][0], ci[t][1], av.y, bv.x);
                        }
                    }
                }
            }
        }
        buf = (buf + 1 == ST) ? 0 : buf + 1;
        nbuf = (nbuf + 1 == ST) ? 0 : nbuf + 1;
    }
    if (warp_active) {
        const int row = m0 + warp * 8 + lr;
        if (row < a.M) {
            const cd* Cin = mat_ptr(a.Cin, b);
            cd* Cout = mat_ptr(a.Cout, b);
            const cd* rs = a.rowscale ? a.rowscale + (long long)(b / a.rs_group) * a.rs_stride : (const cd*)0;
            const cd* cs = a.colscale ? a.colscale + (long long)(b / a.cs_group) * a.cs_stride : (const cd*)0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (t < nt) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int col = n0 + t * 8 + 2 * lk + h;
                        if (col < a.N)
                            Cout[(long long)row * a.Cout.ld + col] = zgemm_epilogue(a, Cin, rs, cs, row, col, mk(cr[t][h], ci[t][h]));
                    }
                }
            }
        }
    }
#endif
}
#ifndef KH_HOST_EMU
// CNT (8-row strip, 8-column tile) units of one warp over a full K chunk.  The fragment addresses are 32-bit shared-memory byte
// addresses kept in registers (stage base + per-unit offset, the k-step as an immediate): the generic-pointer form recomputed
// (strip * 8 + lr) * LDA + ... and the shared window base for every load (2-3 dependent IMADs in front of each LDS).  Fragments of
// G units are in flight at once (G = 2: 8 DMMAs = 128+ cycles of tensor work cover the LDS latency, and the registers that four
// units' fragments took now hold the addresses).
__device__ __forceinline__ cd kh_lds_cd(unsigned addr, int imm) {
    cd v;
    // volatile: keeps the load behind the chunk's barrier in program order (a plain asm without a memory clobber may be moved
    // across __syncthreads by the compiler; ptxas still schedules it freely among the DMMAs)
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + (unsigned)imm));
    return v;
}
template <int CNT, int MAXU, int LDB>
__device__ __forceinline__ void zgemm_mma_units(double (&cr)[MAXU][2], double (&ci)[MAXU][2], unsigned sbase,
                                                const unsigned (&aoff)[MAXU], const unsigned (&boff)[MAXU], int ks) {
    constexpr int G = KH_ZG_GROUP < CNT ? KH_ZG_GROUP : CNT;
#pragma unroll
    for (int kk = 0; kk < ZG_BK / 4; ++kk) {
        if (kk < ks) {
#pragma unroll
            for (int g0 = 0; g0 < CNT; g0 += G) {
                cd av[G], bv[G];
#pragma unroll
                for (int t = 0; t < G; ++t) if (g0 + t < CNT) { av[t] = kh_lds_cd(sbase + aoff[g0 + t], kk * 4 * 16); bv[t] = kh_lds_cd(sbase + boff[g0 + t], kk * 4 * LDB * 16); }
#pragma unroll
                for (int t = 0; t < G; ++t) if (g0 + t < CNT) { kh_dmma(cr[g0 + t][0], cr[g0 + t][1], av[t].x, bv[t].x); kh_dmma(ci[g0 + t][0], ci[g0 + t][1], av[t].x, bv[t].y); }
#pragma unroll
                for (int t = 0; t < G; ++t) if (g0 + t < CNT) { kh_dmma(cr[g0 + t][0], cr[g0 + t][1], -av[t].y, bv[t].y); kh_dmma(ci[g0 + t][0], ci[g0 + t][1], av[t].y, bv[t].x); }
            }
        }
    }
}
#endif

// Unit-balanced variant: the CTA tile is (8 NW) x (8 NT) as above and NW warps stage it, but NW + 1 warps compute, and the
// (strip, tile) units of the tile are dealt out evenly over them instead of one 8-row strip per warp.  With NW = 7 a CTA has
// 8 compute warps = 2 per scheduler, whereas 7 strip-warps leave the 4th scheduler of an SM with half the DMMA work of the
// others (and ragged tiles idle whole warps): at n = 98 the busiest scheduler drops from 52 to 44 of a matrix's 169 units.
template <int NW, int NT, int ST, int NC = NW + 1>
KH_DEV void zgemm_body_u(const Cta& c, const zgemm_args& a) {
#ifdef KH_HOST_EMU
    zgemm_body_t<NW, NT>(c, a);
#else
    static_assert(NW == NT, "staging strides assume a square CTA tile");
    constexpr int BM = 8 * NW, BN = 8 * NT, LDB = BN + 2, STAGE = BM * ZG_LDA + ZG_BK * LDB;
    constexpr int MAXU = (NW * NT + NC - 1) / NC;
    const int tiles_n = (a.N + BN - 1) / BN, tiles = ((a.M + BM - 1) / BM) * tiles_n;
    const int b = c.bx / tiles, tile = c.bx - b * tiles;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const cd* A = mat_ptr(a.A, b);
    const cd* B = mat_ptr(a.B, b);
    cd* sm = (cd*)KH_SMEM(c);
    const int tid = c.tid, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int nk = (a.K + ZG_BK - 1) / ZG_BK;
    const int nt = min(NT, (a.N - n0 + 7) >> 3), ns = min(NW, (a.M - m0 + 7) >> 3);
    const int U = ns * nt, cnt = U / NC + (warp < U % NC ? 1 : 0), ustart = warp * (U / NC) + min(warp, U % NC);
    unsigned aoff[MAXU], boff[MAXU];                   // byte offsets of the unit's fragments inside a stage
#pragma unroll
    for (int j = 0; j < MAXU; ++j) {
        const int u = ustart + (j < cnt ? j : 0), strip = u / nt, tl = u - strip * nt;
        aoff[j] = (unsigned)(((strip * 8 + lr) * ZG_LDA + lk) * (int)sizeof(cd));
        boff[j] = (unsigned)((BM * ZG_LDA + lk * LDB + tl * 8 + lr) * (int)sizeof(cd));
    }
    const unsigned sm_u32 = (unsigned)__cvta_generic_to_shared(sm);
    const bool stager = warp < NW;
    const int am = a.transA ? tid % BM : tid >> 4, ak = a.transA ? tid / BM : tid & 15;
    const int bk = tid / BN, bn = tid - bk * BN;
    const cd* asrc = a.transA ? A + (long long)ak * a.A.ld + m0 + am : A + (long long)(m0 + am) * a.A.ld + ak;
    const cd* bsrc = B + (long long)bk * a.B.ld + n0 + bn;
    const int adst = am * ZG_LDA + ak, bdst = BM * ZG_LDA + bk * LDB + bn;
    const bool bn_ok = (n0 + bn) < a.N;
    const long long a_pass = a.transA ? 4LL * a.A.ld : 2LL * NW * a.A.ld, b_pass = 4LL * a.B.ld;
    const long long a_chunk = a.transA ? (long long)ZG_BK * a.A.ld : ZG_BK, b_chunk = (long long)ZG_BK * a.B.ld;
    auto stage = [&](int buf, int kc) {
        cd* s = sm + buf * STAGE;
        const int k0 = kc * ZG_BK;
        if (!a.transA) {
            const bool kok = (k0 + ak) < a.K;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = kok && (m0 + am + 2 * NW * i) < a.M;
                kh_cp_async16(s + adst + 2 * NW * i * ZG_LDA, ok ? asrc + i * a_pass : A, ok);
            }
        } else {
            const bool mok = (m0 + am) < a.M;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = mok && (k0 + ak + 4 * i) < a.K;
                kh_cp_async16(s + adst + 4 * i, ok ? asrc + i * a_pass : A, ok);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = bn_ok && (k0 + bk + 4 * i) < a.K;
            kh_cp_async16(s + bdst + 4 * i * LDB, ok ? bsrc + i * b_pass : B, ok);
        }
        asrc += a_chunk; bsrc += b_chunk;
    };
    double cr[MAXU][2], ci[MAXU][2];
#pragma unroll
    for (int t = 0; t < MAXU; ++t) { cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0; }
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) { if (s < nk && stager) stage(s, s); kh_cp_async_commit(); }
    int buf = 0, nbuf = ST - 1;
    for (int kc = 0; kc < nk; ++kc) {
        kh_cp_async_wait<ST - 2>();
        __syncthreads();
        if (kc + ST - 1 < nk && stager) stage(nbuf, kc + ST - 1);
        kh_cp_async_commit();
        const unsigned sb = sm_u32 + (unsigned)(buf * STAGE * (int)sizeof(cd));
        const int ks = min(ZG_BK / 4, (a.K - kc * ZG_BK + 3) >> 2);
        if (cnt == MAXU) zgemm_mma_units<MAXU, MAXU, LDB>(cr, ci, sb, aoff, boff, ks);
        else if (cnt == MAXU - 1) zgemm_mma_units<MAXU - 1, MAXU, LDB>(cr, ci, sb, aoff, boff, ks);
        else if (cnt == MAXU - 2) zgemm_mma_units<MAXU - 2, MAXU, LDB>(cr, ci, sb, aoff, boff, ks);
        else if (cnt == MAXU - 3) zgemm_mma_units<(MAXU > 3 ? MAXU - 3 : 1), MAXU, LDB>(cr, ci, sb, aoff, boff, ks);
        else if (cnt > 0) {
            for (int kk = 0; kk < ks; ++kk) {
#pragma unroll
                for (int j = 0; j < MAXU; ++j) {
                    if (j < cnt) {
                        const cd av = kh_lds_cd(sb + aoff[j] + (unsigned)(kk * 4 * (int)sizeof(cd)), 0), bv = kh_lds_cd(sb + boff[j] + (unsigned)(kk * 4 * LDB * (int)sizeof(cd)), 0);
                        kh_dmma(cr[j][0], cr[j][1], av.x, bv.x); kh_dmma(ci[j][0], ci[j][1], av.x, bv.y);
                        kh_dmma(cr[j][0], cr[j][1], -av.y, bv.y); kh_dmma(ci[j][0], ci[j][1], av.y, bv.x);
                    }
                }
            }
        }
        buf = (buf + 1 == ST) ? 0 : buf + 1;
        nbuf = (nbuf + 1 == ST) ? 0 : nbuf + 1;
    }
    const cd* Cin = mat_ptr(a.Cin, b);
    cd* Cout = mat_ptr(a.Cout, b);
    const cd* rs = a.rowscale ? a.rowscale + (long long)(b / a.rs_group) * a.rs_stride : (const cd*)0;
    const cd* cs = a.colscale ? a.colscale + (long long)(b / a.cs_group) * a.cs_stride : (const cd*)0;
#pragma unroll
    for (int j = 0; j < MAXU; ++j) {
        if (j < cnt) {
            const int u = ustart + j, strip = u / nt, tl = u - strip * nt;
            const int row = m0 + strip * 8 + lr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + tl * 8 + 2 * lk + h;
                if (row < a.M && col < a.N)
                    Cout[(long long)row * a.Cout.ld + col] = zgemm_epilogue(a, Cin, rs, cs, row, col, mk(cr[j][h], ci[j][h]));
            }
        }
    }
#endif
}
KH_DEV void zgemm56u3_body(const Cta& c, const zgemm_args& a) { zgemm_body_u<7, 7, 3>(c, a); }
KH_DEV void zgemm64p3_body(const Cta& c, const zgemm_args& a) { zgemm_body_p<8, 8, 3>(c, a); }

static inline long long zgemm_padded(int M, int N, int T) { return (long long)((M + T - 1) / T) * T * ((N + T - 1) / T) * T; }
static inline size_t zgemm_smem(int T, int ST) { return (size_t)ST * (T * ZG_LDA + ZG_BK * (T + 2)) * sizeof(cd); }
// Shapes that waste less padding on 56x56 tiles (n = 50: one tile, n = 98: 2x2 tiles) use the unit-balanced kernel, the rest
// the 64x64 strip kernel.  Measured and dropped in round 1: a strip-per-warp 56x56 pipeline, the first-generation double-buffer
// kernels, a 104x104 one-CTA-per-matrix tile (13 warps need > 128 registers for the 13 accumulator tiles), a 3-CTA/SM 2-stage
// variant (no gain), and a persistent grid whose cp.async pipeline runs across tile boundaries (7 % slower).  Round 2: tiles staged
// by TMA bulk copies (one per tile row) with full / empty mbarriers instead of cp.async + one CTA barrier: 30 % slower
// (profiles/r02_experiments.md).
static inline int zgemm_launch(kh_stream_t st, int batch, const zgemm_args& a) {
    if (batch <= 0 || a.M <= 0 || a.N <= 0) return 0;
    const double work = 8.0 * a.M * a.N * a.K * batch;
    // profiler name: products with a handful of columns (the flux columns of the last star product) are matrix-vector work,
    // bound by reading A from HBM, and are reported apart from the tensor-bound GEMMs
    const char* name = a.N <= 8 ? "zgemv" : "zgemm";
    // (56x56 only where it saves at least 5 % of the padded work: at equal padding the 64x64 strip kernel is the faster one,
    //  n = 450: 465 -> 491 solves/s)
    const bool t56 = zgemm_padded(a.M, a.N, 56) * 100 < zgemm_padded(a.M, a.N, 64) * 95;
    const unsigned g56 = (unsigned)batch * ((a.M + 55) / 56) * ((a.N + 55) / 56), g64 = (unsigned)batch * ((a.M + 63) / 64) * ((a.N + 63) / 64);
    if (t56) return kh_launch<zgemm_args, zgemm56u3_body, 256, 2>(dim3(g56), 256, zgemm_smem(56, 3), st, a, name, work);
    return kh_launch<zgemm_args, zgemm64p3_body, 256, 2>(dim3(g64), 256, zgemm_smem(64, 3), st, a, name, work);
}

// convenience builder: plain C = alpha*A*B (+ beta*Cin) on [batch, n, n] row-major stacks
static inline zgemm_args zgemm_make(int M, int N, int K, MatRef A, MatRef B, MatRef Cout, double alpha = 1.0) {
    zgemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.A = A; g.B = B; g.Cout = Cout;
    g.Cin = mref(nullptr, 0, 0);
    g.alpha = alpha; g.beta = 0.0; g.diag = 0.0;
    g.rs_group = g.cs_group = 1;
    return g;
}
