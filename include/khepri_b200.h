/* khepri_b200 -- C ABI of the B200-native batched RCWA solve engine.
 *
 * The reference (Kaeryv/Khepri) is pure Python and has no FFI layer; its boundary is the class
 * surface of khepri.crystal.Crystal (khepri/crystal.py:42-400).  This header is the C-ABI a
 * binding for that path would bind: plain pointers and sizes, no torch types, no exceptions.
 * Every pointer suffixed _dev is DEVICE memory owned by the caller (the Python host side allocates
 * it through torch); all work is enqueued on `stream` (a cudaStream_t passed as void*) and is
 * stream-ordered.  Functions return 0 on success, a negative KH_E* code for argument errors and a
 * positive cudaError_t for CUDA failures; kh_last_error() gives a thread-local message.
 *
 * complex128 values are interleaved (re, im) doubles, matrices are row-major and unpadded, i.e.
 * exactly numpy's / torch's C-contiguous complex128 layout.
 *
 * Each entry point cites the reference code it replaces.
 */
#ifndef KHEPRI_B200_H
#define KHEPRI_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KH_ABI_VERSION 2
#define KH_EINVAL (-1)
#define KH_ENOMEM (-2)   /* workspace too small */
#define KH_ESTATE (-3)

/* layer kinds = khepri/layer.py:19-24 (Formulation) */
#define KH_LAYER_UNIFORM 0
#define KH_LAYER_PIXMAP 1          /* FFT and ANALYTICAL formulations both arrive here as a convolution matrix */
#define KH_LAYER_HALF_INC 3
#define KH_LAYER_HALF_TRN 4
#define KH_LAYER_EXTENDED 5        /* khepri/extension.py:66-112: base layer solved at N shifted k-points, scattered into the moire basis */

typedef struct {
    int kind;
    double eps_re, eps_im;         /* UNIFORM / HALF_*: permittivity */
    double depth;
    const void* C_dev;             /* PIXMAP: convolution matrix [N][N] c128 (kh_convmat output) */
    const void* IC_dev;            /* PIXMAP: its inverse [N][N] c128 (kh_zinv_batched output); layer.py:158 */
    int retain;                    /* keep W, V, lambda for field reconstruction (Layer.fields) */
    int ext_base;                  /* EXTENDED: index of the base layer in the plan's layer table */
    int ext_mode;                  /* EXTENDED: 0 -> joint = shift*Nb + r ; 1 -> joint = r*Nb + shift (extension.py:35-38) */
} kh_layer_desc;

typedef struct kh_plan kh_plan;

/* optional outputs of kh_solve_batch (any pointer may be NULL) */
typedef struct {
    void* Stot_dev;                /* [B][2][2][n][n] c128   (Crystal.Stot, crystal.py:203-206) */
    double* RT_dev;                /* [B][2]  (R, T)         (Crystal.poynting_flux_end, crystal.py:363-396) */
    double* orders_dev;            /* [B][2][N] per-order fluxes (only_total=False) */
    void* prefix_dev;              /* [B][Ls][2][2][n][n] forward partial products  (layer.py:41-47)  */
    void* suffix_dev;              /* [B][Ls][2][2][n][n] reverse partial products  (layer.py:49-59)  */
    void* W_dev;                   /* [B][n_layers][n][n] layer eigenvectors, E part (Layer.W), per layer-table index */
    void* V_dev;                   /* [B][n_layers][n][n] layer eigenvectors, H part (Layer.V)                        */
    void* L_dev;                   /* [B][n_layers][n]    layer eigenvalues lambda (Layer.L)                          */
    int* info_dev;                 /* [B] 0 = ok; bit0 eigensolver did not converge, bit1 singular pivot, bit2 doubling bound exceeded,
                                      bit3 doubling method: a self star product was ill conditioned (re-solve with KH_METHOD_EIG) */
} kh_outputs;

int kh_abi_version(void);
const char* kh_last_error(void);

/* ---- convolution matrix ------------------------------------------------------------------- */
/* khepri/tools.py:33-56 convolution_matrix(): pruned DFT of L pixmaps [L][Nx][Ny] (f64, or c128 when
 * is_complex) + Toeplitz gather -> C_dev [L][N][N] c128, N = P*Q.  F_dev (optional) receives the
 * compact coefficient table [L][2P-1][2Q-1]. */
size_t kh_convmat_work_bytes(int L, int Nx, int Ny, int P, int Q);
int kh_convmat(int L, int Nx, int Ny, int is_complex, const void* pix_dev, int P, int Q,
               void* C_dev, void* F_dev, void* work_dev, size_t work_bytes, void* stream);
/* khepri/tools.py:38-56 convolution_matrix_fourier(): pure-index gather from a full shifted
 * coefficient array F_dev [Nx][Ny] c128 (bit exact).  err_dev: int, set to 1 on an out-of-range index. */
int kh_toeplitz_gather(const void* F_dev, int Nx, int Ny, int P, int Q, void* C_dev, int* err_dev, void* stream);

/* ---- batched dense complex128 primitives (numpy.linalg call sites) ---------- */
/* C[b] = alpha * op(A[b]) * B[b]; row-major, strides in complex elements (stride 0 broadcasts). */
int kh_zgemm_batched(int batch, int M, int N, int K, int transA,
                     const void* A_dev, int lda, long long strideA,
                     const void* B_dev, int ldb, long long strideB,
                     void* C_dev, int ldc, long long strideC, double alpha, void* stream);
/* numpy.linalg.inv / solve: Ainv[b] = A[b]^-1 (n x n, contiguous stacks); info_dev [batch] optional.
   Matrices beyond shared memory (kh_zinv_work_bytes > 0) use the blocked Gauss-Jordan + DMMA GEMM
   variant and need that much DEVICE work space; work_dev may be NULL when the size is 0. */
size_t kh_zinv_work_bytes(int batch, int n);
int kh_zinv_batched(int batch, int n, const void* A_dev, void* Ainv_dev, int* info_dev,
                    void* work_dev, size_t work_bytes, void* stream);
/* numpy.linalg.eig (alternative.py:172): w[b] eigenvalues [n], W[b] right eigenvectors [n][n] (columns) */
size_t kh_zgeev_work_bytes(int batch, int n);
int kh_zgeev_batched(int batch, int n, const void* A_dev, void* w_dev, void* W_dev,
                     void* work_dev, size_t work_bytes, int* info_dev, void* stream);

/* ---- plan = Crystal geometry (crystal.py:42-164) ------------------------------------------- */
/* g_dev: [2][N] f64 reciprocal vectors of the expansion (expansion.py:36-40), DEVICE memory that
 * must outlive the plan.  layers: table of distinct layers (half spaces included explicitly, as
 * Crystal.set_device adds "Sref"/"Strans"); stack: indices into it, incidence side first.
 * For EXTENDED layers glhs_dev / grhs_dev [2][Nb] are the g-vectors of the two base lattices whose
 * Minkowski sum (expansion.py:55-73, joint index = i_lhs*Nb + i_rhs) is g_dev; a layer with
 * ext_mode 1 lives on the lhs lattice and is shifted by the rhs vectors, ext_mode 0 the other way
 * round (extension.py:66-80).  Pass 0/NULL when there are none. */
int kh_plan_create(kh_plan** plan, int P, int Q, const double* g_dev,
                   double epsi_re, double epsi_im, double epse_re, double epse_im,
                   int n_layers, const kh_layer_desc* layers, int n_stack, const int* stack,
                   int Nb, const double* glhs_dev, const double* grhs_dev);
void kh_plan_destroy(kh_plan* plan);

/* How patterned layers get their S-matrix when no eigenspace has to be retained (flux / Stot outputs).
 *   KH_METHOD_EIG       eigen-decomposition of Omega^2 = P Q (alternative.py:158-195, the Crystal path's algorithm).
 *   KH_METHOD_DOUBLING  the reference's legacy algorithm (khepri/tmat/scattering.py:25-51): transfer matrix of a thin
 *                       slice (power series of exp) -> S-matrix (matrix_s, tmat/matrices.py:167-176) -> self star
 *                       products; GEMMs and inverses only.  kappa bounds k0 * sqrt(rho(Omega^2)) over the batch, i.e.
 *                       kappa * depth bounds the largest |lambda k0 d| of a layer; the layer is cut into 2^s slices with
 *                       kappa * depth / 2^s <= theta_slice.  If the bound turns out too small for a solve (checked on the
 *                       device against ||Omega^2||_1), bit 2 of its info is set.
 * Solves that retain eigenspaces (KH_WANT_FIELDS) always use KH_METHOD_EIG.
 * The setting is part of the plan's state: call it before kh_solve_workspace_bytes / kh_solve_batch of the batch it applies to (a plan
 * is thread-compatible, not thread-safe: one thread at a time per handle). */
#define KH_METHOD_EIG 0
#define KH_METHOD_DOUBLING 1
int kh_plan_set_method(kh_plan* plan, int method, double kappa, double theta_slice);

/* ---- batched solve = Crystal.solve + poynting_flux_end over B (wavelength, k-point) pairs --- */
#define KH_WANT_STOT 1
#define KH_WANT_FLUX 2
#define KH_WANT_FIELDS 4           /* prefix/suffix products + retained eigenspaces */
size_t kh_solve_workspace_bytes(const kh_plan* plan, int chunk, int flags);
/* wl_dev [B] f64; kp_dev [B][2] c128 (crystal.py:345-360 set_source); pol_dev [B][2] c128 = (te, tm).
 * The batch is processed in chunks sized to ws_bytes (at least one solve must fit). */
int kh_solve_batch(kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* pol_dev,
                   const kh_outputs* out, void* ws_dev, size_t ws_bytes, void* stream);

/* S = SA (*) SB on stacks of full S-matrices [B][2][2][n][n] (alternative.py:19-30 redheffer_product;
 * examples/crystal_api/woodpile.py:85 does cl.Stot = redheffer_product(cl.Stot, cl.Stot)). */
size_t kh_star_workspace_bytes(int B, int n);
int kh_star_batch(int B, int n, const void* SA_dev, const void* SB_dev, void* SO_dev,
                  void* ws_dev, size_t ws_bytes, void* stream);
/* poynting_flux_end on caller-provided S-matrices (after the caller modified Stot) */
int kh_flux_batch(const kh_plan* plan, int B, const void* Stot_dev, const double* wl_dev, const void* kp_dev,
                  const void* pol_dev, double* RT_dev, double* orders_dev, void* stream);

/* ---- field reconstruction (crystal.py:234-343, fields.py, fourier.py:136-142) ---------------- */
/* E,H at arbitrary in-plane points (x[p], y[p]), p < npts, and depths z[nz] for every solve of the
 * batch, from the KH_WANT_FIELDS outputs of kh_solve_batch (+ Stot).  inc_dev [B][2][n] c128 are the
 * incident (E, H) Fourier vectors (Crystal.get_source_as_field_vectors).  z_host / zpos_host are HOST
 * arrays: depths, and the Ls+1 interface positions (Crystal.stack_positions, +-inf at the ends).
 * F_dev: [B][nz][6][npts] c128 = (Ex,Ey,Ez,Hx,Hy,Hz). */
size_t kh_fields_workspace_bytes(const kh_plan* plan, int B, int npts, int nz);
int kh_fields_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                    const kh_outputs* solved, const double* x_dev, const double* y_dev, int npts,
                    const double* z_host, int nz, const double* zpos_host, void* F_dev,
                    void* ws_dev, size_t ws_bytes, void* stream);

/* Same on a rectangular grid x = xs[ix], y = ys[iy] (fields_volume with meshgrid coordinates):
 * F_dev [B][nz][6][ny][nx].  Separable inverse transform, 8 (N nx + Q nx ny) flops per map instead of
 * 8 N nx ny.  Requires a lattice whose reciprocal vector b1 has no y component (ky depends on the q
 * index only: square / rectangular lattices) and Q <= 16; the caller checks the lattice
 * (khepri_b200.Crystal does) and falls back to kh_fields_batch otherwise. */
size_t kh_fields_grid_workspace_bytes(const kh_plan* plan, int B, int nx, int ny, int nz);
int kh_fields_grid_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                         const kh_outputs* solved, const double* xs_dev, int nx, const double* ys_dev, int ny,
                         const double* z_host, int nz, const double* zpos_host, void* F_dev,
                         void* ws_dev, size_t ws_bytes, void* stream);

/* Fourier fields only (crystal.py:234-277 _fourier_fields; fields_coords_xy(..., return_fourier=True), :326-327):
 * S_dev [B][nz][6][N] c128 = (sx, sy, sz, ux, uy, uz) per depth, the coefficient vectors the inverse transform of
 * kh_fields_batch consumes.  Workspace: kh_fields_workspace_bytes(plan, B, 1, nz). */
int kh_fields_fourier_batch(const kh_plan* plan, int B, const double* wl_dev, const void* kp_dev, const void* inc_dev,
                            const kh_outputs* solved, const double* z_host, int nz, const double* zpos_host, void* S_dev,
                            void* ws_dev, size_t ws_bytes, void* stream);

/* fourier.idft (fourier.py:136-142) as a standalone operator on scattered points:
 * out_dev [M][npts] c128 = sum_g s_dev[m][g] exp(i (kx[g] x[p] + ky[g] y[p])); kx_dev, ky_dev [N] c128 (callers pass k0 * Kx,
 * k0 * Ky as crystal.py:329 does), x_dev, y_dev [npts] f64, s_dev [M][N] c128. */
size_t kh_idft_work_bytes(int N, int npts);
int kh_idft_batch(int M, int N, int npts, const void* kx_dev, const void* ky_dev, const double* x_dev, const double* y_dev,
                  const void* s_dev, void* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---- Brillouin-zone-integration source (beams.py:164-191, amplitudes_from_fields) ---------------- */
/* amp_dev [B][N][4] c128 = scale * sum_p fields_dev[p][c] exp(-i ((kp[b] + g) . r_p)), c = (Ex, Ey, Hx, Hy);
 * kp_dev [B][2] c128, g_dev [2][N] f64, x_dev / y_dev [npts] f64, fields_dev [npts][4] c128. */
size_t kh_beam_amplitudes_work_bytes(int B, int N, int npts);
int kh_beam_amplitudes(int B, int N, int npts, const void* kp_dev, const double* g_dev, const double* x_dev, const double* y_dev,
                       const void* fields_dev, double scale, void* amp_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---- measurement helper --------------------------------------------------------------------- */
/* FP64 peak probe on the current device (registers only): mode 0 = DFMA stream, 1 = DMMA m8n8k4
 * stream.  Synchronous; returns TFLOP/s.  Used by bench.py for the roofline denominator that
 * MEASURED_PEAKS.json does not carry. */
int kh_fp64_peak(int mode, int iters, int blocks, double* scratch_dev, double* tflops_out);
/* number of kernels this library has launched so far in this process */
long long kh_launch_count(void);
/* per-kernel CUDA-event timing: begin() arms it, end() synchronises and writes one line per kernel
 * class into buf: "name count total_ms total_algorithmic_flops". */
int kh_profile_begin(void);
int kh_profile_end(char* buf, size_t len);

#ifdef __cplusplus
}
#endif
#endif
