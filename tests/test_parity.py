"""Parity of the kernel path (through the C ABI) with the oracle and with the golden vectors produced
by the unmodified reference.  Tolerance from BASELINE.json's north_star: R/T (and S blocks) within
1e-9 relative in complex128; convolution-matrix indexing bit exact."""
import numpy as np
import pytest

from oracle import rcwa_oracle as orc
from tests import cases
from tests.util import BACKENDS, METHODS, build_crystal, engine, gold, sweep_sources

RTOL = 1e-9


def rt_close(got, want, tol=RTOL):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), np.abs(got - want).max()


# ----------------------------------------------------------------------------- dense primitives
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 3), (50, 50, 50), (64, 64, 16), (65, 130, 33), (98, 98, 98), (98, 16, 98), (9, 98, 50), (104, 104, 8), (57, 57, 1), (112, 50, 19)])
def test_zgemm(backend, shape):
    eng = engine(backend)
    rng = np.random.default_rng(1)
    M, N, K = shape
    A = rng.standard_normal((3, M, K)) + 1j * rng.standard_normal((3, M, K))
    B = rng.standard_normal((3, K, N)) + 1j * rng.standard_normal((3, K, N))
    C = eng.zgemm(A, B).cpu().numpy()
    assert np.abs(C - A @ B).max() <= 1e-13 * K * np.abs(A).max() * np.abs(B).max()
    Ct = eng.zgemm(np.ascontiguousarray(A.transpose(0, 2, 1)), B, transA=True, alpha=-2.0).cpu().numpy()
    assert np.abs(Ct + 2 * (A @ B)).max() <= 2e-13 * K * np.abs(A).max() * np.abs(B).max()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 18, 24, 33, 50, 64, 65, 72, 98, 100, 104])
def test_zinv(backend, n):
    eng = engine(backend)
    rng = np.random.default_rng(2)
    A = rng.standard_normal((4, n, n)) + 1j * rng.standard_normal((4, n, n))
    A[1] = np.roll(A[1], 1, axis=0) * 1e-3 + np.eye(n)[::-1]           # forces row interchanges
    Ai, info = eng.zinv(A, return_info=True)
    assert int(info.max().item()) == 0
    Ai = Ai.cpu().numpy()
    ref = np.linalg.inv(A)
    assert np.abs(Ai - ref).max() <= 1e-10 * np.abs(ref).max()
    _, info = eng.zinv(np.zeros((1, n, n), dtype=complex), return_info=True)
    assert int(info.max().item()) != 0                                  # singular -> flagged


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n", [2, 18, 50, 98])
def test_zgeev(backend, n):
    eng = engine(backend)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    A[1] *= np.logspace(-3, 3, n)[None, :]                               # badly scaled: exercises balancing
    A[2] = np.triu(A[2])                                                 # already triangular
    w, W, info = eng.zgeev(A)
    assert int(info.max().item()) == 0
    w, W = w.cpu().numpy(), W.cpu().numpy()
    for b in range(3):
        res = np.abs(A[b] @ W[b] - W[b] * w[b][None, :]).max()
        assert res <= 1e-11 * np.abs(A[b]).max() * np.abs(W[b]).max(), res
        assert np.abs(np.sort_complex(np.linalg.eigvals(A[b])) - np.sort_complex(w[b])).max() <= 1e-8 * np.abs(w[b]).max()


# ----------------------------------------------------------------------------- convolution matrix
@pytest.mark.parametrize("backend", BACKENDS)
def test_toeplitz_gather_bit_exact(backend):
    eng = engine(backend)
    g = gold("toeplitz")
    C = eng.toeplitz_gather(g["coded"], tuple(g["pw"])).cpu().numpy()
    assert np.array_equal(C, g["C"])
    with pytest.raises(IndexError):
        eng.toeplitz_gather(np.zeros((4, 4), dtype=complex), (5, 5))


@pytest.mark.parametrize("backend", BACKENDS)
def test_convmat_vs_reference(backend):
    eng = engine(backend)
    g = gold("convmat")
    pm = cases.disc_pixmap((96, 64), 2.25, (0.05, -0.1), 0.3, 6.0)
    C = eng.convmat(pm, (5, 3))[0].cpu().numpy()
    assert np.abs(C - g["C"]).max() <= 1e-14 * np.abs(g["C"]).max()
    pm2 = cases.disc_pixmap((128, 128), 12, (0, 0), 0.4, 1.0)
    C2, F2 = eng.convmat(np.stack([pm2, pm2.T]), (7, 7), return_coefficients=True)
    assert np.abs(C2[0].cpu().numpy() - g["C77"]).max() <= 1e-14 * np.abs(g["C77"]).max()
    # the gather itself is pure indexing: bit exact against the oracle's gather of the same coefficient table
    F = F2[1].cpu().numpy()
    full = np.zeros((26, 26), dtype=complex)
    full[13 - 6:13 + 7, 13 - 6:13 + 7] = F
    assert np.array_equal(C2[1].cpu().numpy(), orc.toeplitz_gather(full, (7, 7)))
    # complex (lossy) pixmap and a 1-D grating (N, 1)
    pmc = pm.astype(complex) * (1 - 0.05j)
    assert np.abs(eng.convmat(pmc, (5, 3))[0].cpu().numpy() - orc.convolution_matrix(pmc, (5, 3))).max() <= 1e-14 * np.abs(pmc).max()
    line = cases.rect_pixmap((200, 1), 1, (0, 0), (0.4, 2), 9.0)
    assert np.abs(eng.convmat(line, (9, 1))[0].cpu().numpy() - orc.convolution_matrix(line, (9, 1))).max() <= 1e-14 * 9


# ----------------------------------------------------------------------------- spectra
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("backend", BACKENDS)
def test_suh03_spectrum_golden(backend, method):
    """README suh03 (configs[0]): 5x5 harmonics, [Scyl, S1, Scyl], 151 frequencies."""
    eng = engine(backend)
    g = gold("suh03")
    st, srcs = cases.case_suh03()
    idx = list(range(151)) if backend == "cuda" else list(range(0, 151, 10)) + [150]
    cl = build_crystal(st, eng, method=method)
    R, T = sweep_sources(cl, [srcs[i] for i in idx])
    rt_close(np.stack([R, T], 1), g["RT"][idx])
    (Rs, Ro), (Ts, To), S = sweep_sources(cl, [srcs[i] for i in g["Sidx"]], only_total=False, return_S=True)
    rt_close(S, g["Stot"], 1e-9)
    rt_close(np.stack([Ro, To], 1), g["orders"][g["Sidx"]])


@pytest.mark.parametrize("backend", BACKENDS)
def test_scalar_api_matches_reference_loop(backend):
    """The reference's own loop: set_source; solve; poynting_flux_end (README.md:55-59)."""
    eng = engine(backend)
    g = gold("suh03")
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng)
    with pytest.raises(AssertionError):
        cl.solve()                                       # "Call set_source before solving."
    with pytest.raises(AssertionError):
        cl.poynting_flux_end()                           # "Call solve first"
    for i in (3, 77):
        cl.set_source(**srcs[i])
        cl.solve()
        rt_close(cl.poynting_flux_end(), g["RT"][i])
    (Rs, Rg), (Ts, Tg) = cl.poynting_flux_end(only_total=False)
    rt_close([Rs, Ts], g["RT"][77])
    rt_close(np.stack([Rg, Tg]), g["orders"][77])
    assert cl.Stot.shape == (2, 2, 50, 50) and cl.depth == pytest.approx(2.2)
    assert cl.stack_positions[0] == -np.inf and cl.stack_positions[-1] == np.inf


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("pw,nk,nwl,tag", [((3, 3), 4, 5, "bzi33"), ((7, 7), 3, 3, "bzi77")])
def test_bzi_stack(backend, pw, nk, nwl, tag, method):
    """configs[1]: 16-layer grating stack, epse=4, explicit k-points of the Brillouin-zone grid."""
    if backend == "emu" and pw == (7, 7):
        pytest.skip("covered on the GPU; too slow in emulation")
    eng = engine(backend)
    st, srcs = cases.case_bzi(pw, nk, nwl)
    cl = build_crystal(st, eng, method=method)
    R, T = sweep_sources(cl, srcs)
    rt_close(np.stack([R, T], 1), gold(tag)["RT"])


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("backend", BACKENDS)
def test_woodpile_and_doubling(backend, method):
    """configs[2] geometry at 5x5 (+ 11x11 on the GPU) incl. Stot (*) Stot (woodpile.py:85)."""
    from khepri_b200.alternative import redheffer_product
    eng = engine(backend)
    g = gold("woodpile55")
    st, srcs = cases.case_woodpile((5, 5), 3, 3)
    cl = build_crystal(st, eng, method=method)
    sel = range(9) if backend == "cuda" else (0, 4)
    for i in sel:
        cl.set_source(**srcs[i])
        cl.solve()
        rt_close(cl.poynting_flux_end(), g["RT"][i])
        cl.Stot = redheffer_product(cl.Stot, cl.Stot, engine=eng)
        rt_close(cl.poynting_flux_end(), g["RT_doubled"][i])
    # the reference's builder (factory.py:3-24) with woodpile.py:36-39 parameters: same pixmaps bit for bit, same fluxes
    from khepri_b200.factory import make_woodpile
    wp = make_woodpile(0.28, 3.6 ** 2, 0.5, 1.414 / 4, (5, 5), (256, 256), engine=eng)
    for name in "ABCD":
        assert np.array_equal(np.asarray(wp.layers[name].epsilon), st["layers"][name][1])
    wp.set_source(**srcs[4])
    wp.solve()
    rt_close(wp.poynting_flux_end(), g["RT"][4])
    if backend == "cuda":
        st, srcs = cases.case_woodpile((11, 11), 2, 2)
        cl = build_crystal(st, eng, method=method)
        R, T = sweep_sources(cl, srcs)
        rt_close(np.stack([R, T], 1), gold("woodpile1111")["RT"])


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("backend", BACKENDS)
def test_oblique_hexagonal_lossy(backend, method):
    eng = engine(backend)
    g = gold("oblique")
    st, srcs = cases.case_oblique()
    cl = build_crystal(st, eng, method=method)
    (Rs, Ro), (Ts, To) = sweep_sources(cl, srcs, only_total=False)
    rt_close(np.stack([Rs, Ts], 1), g["RT"])
    rt_close(np.stack([Ro, To], 1), g["orders"])
    cl.set_source(**srcs[2])
    cl.solve()
    rt_close(cl.poynting_flux_end(), g["RT"][2])


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("which", ["tidy", "mixed", "rect"])
def test_analytical_layers(backend, which, method):
    """Crystal.add_layer_analytical -- host-side analytic island transforms, device Toeplitz gather + inverse,
    then the same batched layer solve as a pixmap layer.  Golden vectors from the unmodified reference."""
    eng = engine(backend)
    g = gold("analytical")
    st, srcs = cases.case_analytical(which)
    cl = build_crystal(st, eng, method=method)
    R, T = sweep_sources(cl, srcs)
    rt_close(np.stack([R, T], 1), g["RT_" + which])
    name = [k for k, v in st["layers"].items() if v[0] == "analytical"][0]
    Cm, ICm = cl.layers[name].convmat_device(eng)
    Cm = Cm.cpu().numpy()
    assert np.abs(Cm - g["C_" + which]).max() <= 1e-14 * np.abs(g["C_" + which]).max()
    assert np.abs(ICm.cpu().numpy() @ Cm - np.eye(Cm.shape[0])).max() <= 1e-11


@pytest.mark.parametrize("backend", BACKENDS)
def test_fresnel_known_answer(backend):
    """The reference's only asserted Crystal-path test (test/integration/test_complex_eps.py)."""
    from khepri_b200 import Crystal
    eng = engine(backend)
    fcases, rfres = cases.case_fresnel()
    g = gold("fresnel")
    R = []
    for st, src in fcases:
        cl = Crystal((1, 1), engine=eng)
        cl.add_layer_uniform("1", st["layers"]["1"][1], st["layers"]["1"][2])
        cl.set_device(["1"])
        cl.set_source(**src)
        cl.solve()
        R.append(cl.poynting_flux_end())
    R = np.array(R)
    np.testing.assert_allclose(rfres, R[:, 0], rtol=1e-7)
    rt_close(R, g["RT"], 1e-12)


@pytest.mark.parametrize("backend", BACKENDS)
def test_star_product_matches_oracle(backend):
    eng = engine(backend)
    rng = np.random.default_rng(5)
    n = 18
    SA = 0.3 * (rng.standard_normal((3, 2, 2, n, n)) + 1j * rng.standard_normal((3, 2, 2, n, n)))
    SB = 0.3 * (rng.standard_normal((3, 2, 2, n, n)) + 1j * rng.standard_normal((3, 2, 2, n, n)))
    SO = eng.star(SA, SB).cpu().numpy()
    for b in range(3):
        ref = orc.star(SA[b], SB[b])
        assert np.abs(SO[b] - ref).max() <= 1e-11 * np.abs(ref).max()
    ident = orc.identity_smatrix(n)
    assert np.abs(eng.star(ident, SA[0]).cpu().numpy() - SA[0]).max() <= 1e-14
    assert np.abs(eng.star(SA[0], ident).cpu().numpy() - SA[0]).max() <= 1e-14


# ----------------------------------------------------------------------------- size-independent properties
@pytest.mark.parametrize("backend", BACKENDS)
def test_energy_conservation_and_batch_invariance(backend):
    """Lossless stacks: R + T = 1; results do not depend on how the batch is chunked or ordered."""
    eng = engine(backend)
    st, srcs = cases.case_suh03()
    nb = 151 if backend == "cuda" else 12
    cl = build_crystal(st, eng)
    wl = np.array([s["wavelength"] for s in srcs[:nb]])
    kx = np.linspace(0, 0.3 * np.pi, nb)
    kps = np.stack([kx, 0 * kx], 1)
    R, T = cl.solve_batch(wl, kps=kps, te=1.0, tm=0.0)
    assert np.abs(R + T - 1).max() < 1e-10
    perm = np.random.default_rng(0).permutation(nb)
    R2, T2 = cl.solve_batch(wl[perm], kps=kps[perm], te=1.0, tm=0.0, chunk=5)
    assert np.array_equal(R2, R[perm]) and np.array_equal(T2, T[perm])
    R0, T0 = cl.solve_batch(np.zeros(0), kps=np.zeros((0, 2)))
    assert R0.shape == (0,) and T0.shape == (0,)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n", [101, 119, 150, 242, 450])
@pytest.mark.parametrize("variant", ["auto", "l2", "blocked"])
def test_zinv_blocked(backend, n, variant, monkeypatch):
    """Inverse of matrices beyond one SM's shared memory: the blocked multi-launch variant (Gauss-Jordan panels + DMMA GEMM
    updates) and, for small batches up to n = 256, the two single-launch variants: one thread-block cluster per matrix with the
    matrix in distributed shared memory ("auto" picks it for this batch) and one CTA per matrix with the working copy in L2."""
    if backend != "cuda" and (n > 150 or variant != "auto"):
        pytest.skip("host emulation: small sizes only, one variant")
    if variant in ("l2", "blocked"):
        monkeypatch.setenv("KH_ZINV_CLUSTER_MAXCTAS", "0")
    if variant == "blocked":
        monkeypatch.setenv("KH_ZINV_L2_MAXBATCH", "0")
    eng = engine(backend)
    rng = np.random.default_rng(12)
    A = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    A[1] = np.roll(A[1], 1, axis=0) * 1e-3 + np.eye(n)[::-1]           # forces row interchanges in every panel
    A[2] = np.eye(n) + 1e-2 * A[2]
    Ai, info = eng.zinv(A, return_info=True)
    assert int(info.max().item()) == 0
    Ai = Ai.cpu().numpy()
    ref = np.linalg.inv(A)
    assert np.abs(Ai - ref).max() <= 1e-10 * np.abs(ref).max()
    _, info = eng.zinv(np.zeros((1, n, n), dtype=complex), return_info=True)
    assert int(info.max().item()) != 0                                  # singular -> flagged


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n", [120, 131, 170, 242, 300, 450])
def test_zgeev_tiled(backend, n):
    """Tiled eigensolver (blocked Hessenberg panels + DMMA GEMM updates, tiled QR sweeps) for n beyond shared memory."""
    if backend != "cuda" and n > 170:
        pytest.skip("host emulation: small sizes only")
    eng = engine(backend)
    rng = np.random.default_rng(13)
    A = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    A[1] *= np.logspace(-3, 3, n)[None, :]                               # badly scaled: exercises balancing
    A[2] = np.triu(A[2], -1)                                             # already Hessenberg: trivial reflectors
    A[2, 40:, :40] = 0                                                   # and reducible
    w, W, info = eng.zgeev(A)
    assert int(info.max().item()) == 0
    w, W = w.cpu().numpy(), W.cpu().numpy()
    for b in range(3):
        res = np.abs(A[b] @ W[b] - W[b] * w[b][None, :]).max()
        assert res <= 1e-11 * np.abs(A[b]).max() * np.abs(W[b]).max(), res
        assert np.abs(np.sort_complex(np.linalg.eigvals(A[b])) - np.sort_complex(w[b])).max() <= 1e-7 * np.abs(w[b]).max()


@pytest.mark.parametrize("backend", BACKENDS)
def test_new_entry_points_edge_cases(backend):
    """Empty batches and argument checks of the entry points added for the tiled / grid / BZI rows."""
    import ctypes as C
    from khepri_b200 import Expansion
    eng = engine(backend)
    lib = eng.lib
    # blocked inverse: work space is mandatory above the shared-memory sizes, and sized by kh_zinv_work_bytes
    assert lib.kh_zinv_work_bytes(4, 50) == 0 and lib.kh_zinv_work_bytes(4, 242) >= 4 * 32 * 242 * 16
    A = eng.to_dev(np.eye(130, dtype=complex)[None], __import__("torch").complex128)
    out = A.clone()
    rc = lib.kh_zinv_batched(1, 130, C.c_void_p(A.data_ptr()), C.c_void_p(out.data_ptr()), None, None, 0, eng.stream())
    assert rc != 0 and b"workspace" in lib.kh_last_error()
    assert np.array_equal(eng.zinv(np.eye(130, dtype=complex)[None]).cpu().numpy()[0], np.eye(130))        # exact on the identity
    # beam amplitudes: empty k-batch, and a one-sample "beam" (a single exponential per harmonic)
    e = Expansion((3, 1))
    amp = eng.beam_amplitudes(np.zeros((0, 2)), e._g_vectors, np.array([0.3]), np.array([0.2]), np.ones((1, 4)), 1.0)
    assert tuple(amp.shape) == (0, 3, 4)
    amp = eng.beam_amplitudes(np.array([[0.5, -0.25]]), e._g_vectors, np.array([0.3]), np.array([0.2]), np.array([[1, 2, 3, 4]], dtype=complex), 0.5).cpu().numpy()
    ref = 0.5 * np.exp(-1j * ((0.5 + e._g_vectors[0]) * 0.3 + (-0.25 + e._g_vectors[1]) * 0.2))[:, None] * np.array([1, 2, 3, 4])[None, :]
    assert np.abs(amp[0] - ref).max() <= 1e-14


# ----------------------------------------------------------------------------- sizes beyond shared memory / full-size properties
@pytest.mark.parametrize("backend", BACKENDS)
def test_large_matrices_take_the_global_memory_paths(backend):
    """n > 118 does not fit in shared memory: inverse and eigensolver run on HBM/L2-resident matrices."""
    eng = engine(backend)
    rng = np.random.default_rng(7)
    n = 130 if backend == "cuda" else 122
    A = rng.standard_normal((2, n, n)) + 1j * rng.standard_normal((2, n, n))
    Ai = eng.zinv(A).cpu().numpy()
    assert np.abs(Ai @ A - np.eye(n)).max() <= 1e-10
    w, W, info = eng.zgeev(A[:1])
    assert int(info.max().item()) == 0
    w, W = w.cpu().numpy(), W.cpu().numpy()
    assert np.abs(A[0] @ W[0] - W[0] * w[0][None, :]).max() <= 1e-10 * np.abs(A[0]).max() * np.abs(W[0]).max()


@pytest.mark.gpu
def test_full_size_bzi_step_properties():
    """BASELINE configs[1] at bench size (16 k-points x 101 wavelengths, 7x7): lossless stack -> R + T = 1;
    the chunked run equals the single-pass run bit for bit; a sample of solves matches the oracle."""
    eng = engine("cuda")
    st = cases.bzi_structure((7, 7))
    kg = cases.bzi_kgrid((64, 64)).reshape(2, -1)
    wls = 1 / np.linspace(0.8, 1.0, 101)
    ks = kg[:, 100:116]
    wl = np.tile(wls, 16)
    kp = np.repeat(ks.T, 101, axis=0)
    cl = build_crystal(st, eng)
    R, T = cl.solve_batch(wl, kps=kp, te=1.0, tm=1.0)
    assert np.isfinite(R).all() and np.abs(R + T - 1).max() < 1e-9
    R2, T2 = cl.solve_batch(wl, kps=kp, te=1.0, tm=1.0, chunk=300)
    assert np.array_equal(R, R2) and np.array_equal(T, T2)
    for i in (0, 517, 1615):
        ref = orc.solve_rt(st, wl[i], 1.0, 1.0, kp=(complex(kp[i, 0]), complex(kp[i, 1])))
        rt_close([R[i], T[i]], ref)


@pytest.mark.gpu
def test_convmat_512_and_batch_of_layers():
    eng = engine("cuda")
    pm = cases.disc_pixmap((512, 512), 12, (0.0, 0.0), 0.4, 1.0)
    pm2 = cases.rect_pixmap((512, 512), 1, (0.1, -0.2), (0.3, 0.6), 9.0)
    C = eng.convmat(np.stack([pm, pm2]), (15, 15)).cpu().numpy()
    for k, p in enumerate((pm, pm2)):
        ref = orc.convolution_matrix(p, (15, 15))
        assert np.abs(C[k] - ref).max() <= 1e-14 * np.abs(ref).max()


# ----------------------------------------------------------------------------- band post-processing
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("tag", ["p5a", "p5b", "p3a", "p3b"])
def test_scattering_eigenvalues(backend, tag):
    """khepri.eigentricks.scattering_eigenvalues (eigentricks.py:29-40) as transfer matrix + batched eigensolver: spectrum
    against the reference's QZ result, and the generalized residual Sl v = w Sr v of every returned pair."""
    from khepri_b200 import eigentricks as et
    from tests.test_oracle_golden import match_spectrum
    eng = engine(backend)
    g = gold("bands")
    S4 = g[tag + "_S"]
    w, v = et.scattering_eigenvalues(S4, engine=eng)
    match_spectrum(w, g[tag + "_w"], 1e-8)
    Sl, Sr = orc.scattering_splitlr(orc.flat_smatrix(S4))
    assert np.abs(Sl @ v - (Sr @ v) * w[None, :]).max() <= 1e-8 * np.abs(w).max() * np.abs(v).max()
    # flat layout + batch axis + determinant
    Sflat = orc.flat_smatrix(S4)
    wb, vb, det = et.scattering_eigenvalues(np.stack([Sflat, Sflat]), dos=True, engine=eng)
    assert wb.shape == (2, w.size) and vb.shape == (2,) + v.shape and det.shape == (2,)
    match_spectrum(wb[1], g[tag + "_w"], 1e-8)
    assert abs(det[0] - g[tag + "_det"]) <= 1e-7 * abs(g[tag + "_det"])
    Slh, Srh = et.scattering_splitlr(Sflat)
    assert np.array_equal(Slh, Sl) and np.array_equal(Srh, Sr)
    bad = S4.copy(); bad[0, 0, 0, 0] = np.nan
    assert et.scattering_eigenvalues(bad, engine=eng) is None
    assert et.band_structure(S4, engine=eng).ndim == 1


# ----------------------------------------------------------------------------- round 2: methods, chunking, cache keys, large bases
@pytest.mark.parametrize("backend", BACKENDS)
def test_doubling_matches_eig_on_deep_and_lossy_layers(backend):
    """The doubling method (slice series + self star products; reference legacy path tmat/scattering.py:25-51) against the
    eigen-decomposition method and the oracle on layers that need several doublings (depth 2.5), lossy pixmaps, oblique
    incidence and a thin layer that needs none."""
    eng = engine(backend)
    pm = cases.disc_pixmap((64, 64), 2.0, (0.1, 0.0), 0.3, 9.0)
    pml = pm * (1 - 0.02j)
    layers = {"deep": ("pixmap", pm, 2.5), "thin": ("pixmap", pml, 0.03), "U": ("uniform", 2.1 - 0.1j, 0.4)}
    st = cases._st((5, 5), layers, ["thin", "U", "deep", "thin"], epsi=1.0, epse=2.25)
    srcs = [dict(wavelength=1.7, te=1.0, tm=0.3, theta=0.0, phi=0.0), dict(wavelength=1.1, te=0.2, tm=1.0, theta=33.0, phi=40.0),
            dict(wavelength=2.9, te=1.0, tm=1.0, theta=70.0, phi=-15.0)]
    ref = np.array([orc.solve_rt(st, s["wavelength"], s["te"], s["tm"], s["theta"], s["phi"]) for s in srcs])
    out = {}
    for method in METHODS:
        cl = build_crystal(st, eng, method=method)
        R, T, S = sweep_sources(cl, srcs, return_S=True)
        rt_close(np.stack([R, T], 1), ref)
        out[method] = S
    rt_close(out["auto"], out["eig"], 1e-9)
    # other slice angles (more / fewer self star products) and the block-by-block Horner path of very long series; the method forced
    # and its conditioning guard switched off, so that the doubling arithmetic itself is what is compared
    import os
    keep = eng.doubling_theta
    try:
        os.environ["KH_DBL_COND_LIMIT"] = "1e300"
        for theta, stepwise in ((2.0, "0"), (5.0, "1"), (12.0, "0")):
            eng.doubling_theta = theta
            os.environ["KH_DBL_STEPWISE"] = stepwise
            cl = build_crystal(st, eng, method="doubling")
            R, T, S = sweep_sources(cl, srcs, return_S=True)
            rt_close(np.stack([R, T], 1), ref)
            rt_close(S, out["eig"], 1e-9)
    finally:
        eng.doubling_theta = keep
        os.environ.pop("KH_DBL_STEPWISE", None)
        os.environ.pop("KH_DBL_COND_LIMIT", None)


@pytest.mark.parametrize("backend", BACKENDS)
def test_doubling_guard_resolves_a_sub_slab_resonance_with_eig(backend):
    """A deep high-contrast grating (depth 2.7: three self star products) hit at a resonance of its quarter slab: D = I - S11^2 of
    that doubling has kappa_1 ~ 1e7 and the doubled S-matrix loses digits (found by tests/test_fuzz_parity.py: 9e-8 in R, T).
    The device raises info bit 3; method "auto" solves that source again with the eigen method, forced "doubling" raises."""
    eng = engine(backend)
    rng = np.random.default_rng(77)
    pm = np.where(rng.random((40, 33)) > 0.5, 12.0, 1.0)
    # (the fuzz trial's pixmap is not reproduced here: scan wavelengths of a similar grating for a flagged source instead)
    layers = {"A": ("pixmap", pm, 2.7), "U": ("uniform", 2.2 - 0.05j, 1.2), "B": ("pixmap", pm[::-1].copy(), 0.68)}
    st = cases._st((7, 1), layers, ["A", "U", "B"], epsi=1.3, epse=2.0)
    wls = np.linspace(1.2, 2.4, 241)
    srcs = [dict(wavelength=float(w), te=0.7, tm=0.6, theta=31.0, phi=12.0) for w in wls]
    cl = build_crystal(st, eng, method="auto")
    before = eng.eig_fallbacks
    R, T = sweep_sources(cl, srcs)
    nfb = eng.eig_fallbacks - before
    assert 0 < nfb < len(srcs) // 4, nfb                       # some sources are flagged, most are not
    cle = build_crystal(st, eng, method="eig")
    Re, Te = sweep_sources(cle, srcs)
    rt_close(np.stack([R, T], 1), np.stack([Re, Te], 1))
    ref = np.array([orc.solve_rt(st, s["wavelength"], s["te"], s["tm"], s["theta"], s["phi"]) for s in srcs[::40]])
    rt_close(np.stack([R, T], 1)[::40], ref)
    with pytest.raises(np.linalg.LinAlgError):
        sweep_sources(build_crystal(st, eng, method="doubling"), srcs)


@pytest.mark.parametrize("backend", BACKENDS)
def test_doubling_flags_an_underestimated_spectrum(backend):
    """The host's bound on the spectrum of Omega^2 is checked on the device: a bound that is far too small raises info bit 2
    (surfaced as LinAlgError) instead of returning a truncated series."""
    eng = engine(backend)
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng, method="doubling")
    plan = cl._get_plan(False)
    keep = (plan.gmax, plan.eps_bound)
    plan.gmax, plan.eps_bound = 1e-3, 0.0
    with pytest.raises(np.linalg.LinAlgError):
        sweep_sources(cl, srcs[:2])
    plan.gmax, plan.eps_bound = keep
    R, T = sweep_sources(cl, srcs[:2])
    rt_close(np.stack([R, T], 1), gold("suh03")["RT"][:2])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("method", METHODS)
def test_whole_batch_is_one_chunk_when_the_workspace_fits(backend, method):
    """kh_solve_batch must not split a batch that fits its workspace (round-1 bug: every batch ran as two half chunks)."""
    eng = engine(backend)
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng, method=method)
    lib = eng.lib
    sweep_sources(cl, srcs[:1])                                            # plan + convolution matrices built
    l0 = lib.kh_launch_count(); sweep_sources(cl, srcs[:1]); one = lib.kh_launch_count() - l0
    l0 = lib.kh_launch_count(); sweep_sources(cl, srcs[:12]); twelve = lib.kh_launch_count() - l0
    assert twelve == one, (one, twelve)
    l0 = lib.kh_launch_count(); sweep_sources(cl, srcs[:12], chunk=6); halves = lib.kh_launch_count() - l0
    assert halves > twelve


@pytest.mark.parametrize("backend", BACKENDS)
def test_pixmap_mutated_in_place_is_seen(backend):
    """The reference recomputes convolution_matrix(layer.epsilon) on every solve (layer.py:157): editing a pixmap in place
    and solving again must give the new structure's result, not a cached one."""
    eng = engine(backend)
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng)
    cl.set_source(**srcs[40])
    cl.solve()
    r0, _ = cl.poynting_flux_end()
    eps = cl.layers["Scyl"].epsilon
    eps[10:50, 10:50] = 9.0
    cl.solve()
    r1, t1 = cl.poynting_flux_end()
    st2 = dict(st); st2["layers"] = dict(st["layers"]); st2["layers"]["Scyl"] = ("pixmap", eps.copy(), st["layers"]["Scyl"][2])
    ref = orc.solve_rt(st2, srcs[40]["wavelength"], 1.0, 0.0)
    rt_close([r1, t1], ref)
    assert abs(r1 - r0) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("pp", [13, 15])
def test_supercell_large_bases_golden(pp, method):
    """C4 large set: direct 13x13 / 15x15 bases (n = 338 / 450; tiled Hessenberg / QR / blocked inverse, or the doubling
    method) against the unmodified reference: R, T and strided subsets of Stot[0,0], Stot[1,0]."""
    eng = engine("cuda")
    g = gold(f"supercell{pp}")
    st, srcs = cases.case_supercell(pp)
    cl = build_crystal(st, eng, method=method)
    R, T, S = sweep_sources(cl, srcs, return_S=True)
    rt_close(np.stack([R, T], 1), g["RT"])
    rt_close(S[:, 0, 0, ::9, ::7], g["S11"])
    rt_close(S[:, 1, 0, ::9, ::7], g["S21"])


# ----------------------------------------------------------------------------- flag protocol of the QR relay sweep (GPU only)
@pytest.mark.gpu
def test_qr_flag_protocol_under_timing_stress():
    """The packed QR sweep hands rotations from warp to warp through release-stored / acquire-loaded progress counters in
    shared memory (kh_zgeev.cuh, relay sweep) -- a pattern racecheck cannot model.  Litmus-style check: random pauses after
    every publication and before every poll (KH_QR_STRESS) must not change ANY bit of the eigenvalues or eigenvectors of
    20 000 matrices (every sweep of every matrix exercises the hand-over ~n times), at the sizes that run 2, 4 and 8 CTAs
    per SM, and the residuals stay at rounding level."""
    import os
    eng = engine("cuda")
    rng = np.random.default_rng(21)
    for n, batch in ((98, 20000), (50, 20000), (18, 4000)):
        A = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
        A[::3] *= np.logspace(-2, 2, n)[None, None, :]
        w0, W0, info0 = eng.zgeev(A)
        w0, W0 = w0.cpu().numpy(), W0.cpu().numpy()
        assert int(info0.max().item()) == 0
        os.environ["KH_QR_STRESS"] = "1"
        try:
            w1, W1, info1 = eng.zgeev(A)
            w1, W1 = w1.cpu().numpy(), W1.cpu().numpy()
        finally:
            os.environ.pop("KH_QR_STRESS", None)
        assert int(info1.max().item()) == 0
        assert np.array_equal(w0, w1) and np.array_equal(W0, W1), n
        for b in (0, batch // 2, batch - 1):
            res = np.abs(A[b] @ W0[b] - W0[b] * w0[b][None, :]).max()
            assert res <= 1e-11 * np.abs(A[b]).max() * np.abs(W0[b]).max(), res


@pytest.mark.gpu
def test_qr_flag_protocol_bail_out_is_reported():
    """A consumer that polls a progress counter more than spin_cap times gives up: with an absurdly small cap (and pauses that
    make producers late) the kernel must terminate and flag the matrices (info = n + 1) instead of hanging or returning garbage
    silently; with the default cap the same input converges."""
    import os
    eng = engine("cuda")
    rng = np.random.default_rng(22)
    A = rng.standard_normal((512, 98, 98)) + 1j * rng.standard_normal((512, 98, 98))
    os.environ["KH_QR_SPIN_CAP"] = "1"
    os.environ["KH_QR_STRESS"] = "1"
    try:
        _, _, info = eng.zgeev(A)
    finally:
        os.environ.pop("KH_QR_SPIN_CAP", None)
        os.environ.pop("KH_QR_STRESS", None)
    info = info.cpu().numpy()
    assert (info == 99).any() and set(np.unique(info)) <= {0, 99}
    _, _, info = eng.zgeev(A)
    assert int(info.max().item()) == 0


@pytest.mark.parametrize("backend", BACKENDS)
def test_random_structures_both_methods_against_the_oracle(backend):
    """Differential test on random structures: pixmaps (smooth, binary, lossy), P != Q and 1-D bases, oblique lattice, thin /
    deep / repeated layers (collapsed runs), random oblique sources -- both layer methods against the oracle (R, T) and
    against each other (full Stot)."""
    eng = engine(backend)
    rng = np.random.default_rng(5)
    ntrial = 24 if backend == "cuda" else 9
    for trial in range(ntrial):
        pw = [(3, 3), (5, 3), (3, 5), (7, 1), (1, 5), (5, 5)][trial % 6]
        res = (int(rng.integers(24, 64)), int(rng.integers(24, 64)))
        pm = rng.uniform(1, 6, size=res)
        if trial % 3 == 0:
            pm = np.where(rng.random(res) > 0.5, 12.0, 1.0)
        if trial % 4 == 1:
            pm = pm * (1 - 0.05j)
        d1, d2 = float(rng.choice([0.02, 0.3, 1.0, 2.7])), float(rng.uniform(0.05, 1.5))
        layers = {"A": ("pixmap", pm, d1), "U": ("uniform", complex(rng.uniform(1, 4), -rng.uniform(0, 0.2)), d2),
                  "B": ("pixmap", pm[::-1].copy(), float(rng.uniform(0.05, 0.8)))}
        lat = np.eye(2) if trial % 5 else 0.8 * np.array([[1, 0], [0.5, np.sqrt(3) / 2]])
        st = cases._st(pw, layers, [["A", "U", "B"], ["A", "A", "U", "B", "B"], ["B", "A"]][trial % 3], lattice=lat,
                       epsi=float(rng.uniform(1, 2)), epse=float(rng.uniform(1, 3)))
        srcs = [dict(wavelength=float(rng.uniform(0.7, 2.5)), te=float(rng.uniform(0, 1)), tm=float(rng.uniform(0.1, 1)),
                     theta=float(rng.uniform(0, 70)), phi=float(rng.uniform(0, 360))) for _ in range(3)]
        ref = np.array([orc.solve_rt(st, s["wavelength"], s["te"], s["tm"], s["theta"], s["phi"]) for s in srcs])
        out = {}
        for method in METHODS:
            cl = build_crystal(st, eng, method=method)
            R, T, S = sweep_sources(cl, srcs, return_S=True)
            rt_close(np.stack([R, T], 1), ref)
            out[method] = S
        rt_close(out["auto"], out["eig"], 1e-9)
