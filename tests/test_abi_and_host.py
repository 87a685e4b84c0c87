"""CPU: the C-ABI library loads and exports every symbol include/khepri_b200.h declares; host-side
logic (expansion, k-vectors, sources, error behaviour) mirrors the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import rcwa_oracle as orc
from tests.util import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "khepri_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kh_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ("kh_convmat", "kh_toeplitz_gather", "kh_zgemm_batched", "kh_zinv_batched", "kh_zgeev_batched", "kh_plan_create",
                 "kh_solve_batch", "kh_star_batch", "kh_flux_batch", "kh_fields_batch", "kh_fields_fourier_batch", "kh_idft_batch"):
        assert must in syms


def test_cuda_library_exports_every_declared_symbol():
    """No compute call is made (there is no GPU here) -- only that the product .so loads and binds."""
    from khepri_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/khepri_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(header_symbols())
    bound = _lib.bind()
    assert bound.kh_abi_version() == 2


def test_no_cpu_fallback_without_cuda():
    import torch
    from khepri_b200 import Crystal, Engine, KhepriError
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(KhepriError):
        Engine()
    cl = Crystal((3, 3))
    cl.add_layer_uniform("U", 2.0, 0.1)
    cl.set_device(["U"])
    cl.set_source(1.0)
    with pytest.raises(KhepriError):
        cl.solve()                       # the product path fails loudly instead of computing on the CPU
    with pytest.raises(KhepriError):
        from khepri_b200 import _lib
        _lib.bind("/nonexistent/libkhepri_b200.so")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "khepri_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|\boracle\.[a-z_]+\(|rcwa_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_expansion_matches_oracle_and_reference_conventions():
    from khepri_b200 import Expansion
    from khepri_b200.expansion import generate_expansion_indices
    with pytest.raises(AssertionError):
        generate_expansion_indices((4, 3))                        # odd pw asserted (expansion.py:10-11)
    lat = 0.9 * np.array([[np.sqrt(3) / 2, 0.5], [np.sqrt(3) / 2, -0.5]])
    e = Expansion((5, 3), lat)
    assert np.array_equal(e.expansion_indices, orc.harmonic_indices((5, 3)))
    assert np.allclose(e.g_vectors, orc.g_vectors((5, 3), lat), rtol=0, atol=0)
    kv = e.k_vectors((0.3 + 0j, -0.2 + 0j), 1.3, 2.25 - 0.1j)
    ko = orc.k_vectors(orc.g_vectors((5, 3), lat), (0.3 + 0j, -0.2 + 0j), 1.3, 2.25 - 0.1j)
    assert np.array_equal(kv, np.stack(ko))
    e1, e2 = Expansion((3, 3)), Expansion((3, 3))
    e1.rotate(0.1)
    e2.rotate(-0.1)
    m = e1 + e2
    assert m.pw == (9, 9) and m.expansion_lhs is e1 and m.expansion_rhs is e2
    assert np.allclose(m.g_vectors, orc.minkowski_sum(e1.g_vectors, e2.g_vectors))


def test_source_and_incident_vector():
    from khepri_b200 import Crystal
    from khepri_b200.alternative import incident
    from khepri_b200.tools import compute_kplanar
    cl = Crystal((3, 3), epsi=1.44)
    cl.set_source(1.3, 0.6, 0.8, theta=20.0, phi=35.0)
    assert np.allclose(cl.kp, orc.kplanar(1.44, 1.3, 20.0, 35.0))
    assert np.allclose(compute_kplanar(1.44, 1.3, 20.0, 35.0), orc.kplanar(1.44, 1.3, 20.0, 35.0))
    kv = (cl.kp[0], cl.kp[1], cl.kzi)
    assert np.allclose(incident((3, 3), 0.6, 0.8, kv), orc.incident_vector((3, 3), 0.6, 0.8, kv))
    cl.set_source(0.9, kp=(0.1, 0.2))
    assert cl.kp == (0.1, 0.2)
    with pytest.raises(NotImplementedError):
        Crystal((3, 3), lattice="kagome")
    from khepri_b200 import Formulation
    cl.add_layer_analytical("a", [{"type": "disc", "params": [0.5, 0.5, 0.2], "epsilon": 2.0}], 1.0, 0.1)     # layer.py:121-129
    assert cl.layers["a"].formulation == Formulation.ANALYTICAL and cl.layers["a"].eps_host == 1.0


def test_set_device_adds_half_spaces_like_the_reference():
    from khepri_b200 import Crystal, Formulation
    cl = Crystal((3, 3), epsi=2.0, epse=3.0)
    cl.add_layer_uniform("U", 2.0, 0.1)
    cl.add_layer_pixmap("P", np.ones((8, 8)) + np.eye(8), 0.2)
    cl.set_device(["U", "P", "U"], [True, False, True])
    assert cl.global_stacking == ["Sref", "U", "P", "U", "Strans"]
    assert cl.stack_retain_mask == [True, True, False, True, True]
    assert cl.layers["Sref"].formulation == Formulation.HALF_SPACE_INC and cl.layers["Sref"].epsilon == 2.0
    assert cl.layers["Strans"].formulation == Formulation.HALF_SPACE_TRN and cl.layers["Strans"].epsilon == 3.0
    assert cl.layers["U"].fields and not cl.layers["P"].fields
    assert cl.depth == pytest.approx(0.4) and not cl.solved and not cl.source_defined


@pytest.mark.gpu
def test_integration_stub_idft_raw_ctypes():
    """The reference-side stub of INTEGRATION.md 2b, verbatim in spirit: raw ctypes on the C ABI (no khepri_b200 host code),
    checked against the unmodified reference's idft output (tests/golden/idft.npz)."""
    import ctypes as C
    import torch
    lib = C.CDLL(os.path.join(ROOT, "khepri_b200", "lib", "libkhepri_b200.so"))
    lib.kh_last_error.restype = C.c_char_p
    lib.kh_idft_work_bytes.restype = C.c_size_t
    lib.kh_idft_work_bytes.argtypes = [C.c_int, C.c_int]
    lib.kh_idft_batch.argtypes = [C.c_int] * 3 + [C.c_void_p] * 7 + [C.c_size_t, C.c_void_p]
    g = np.load(os.path.join(ROOT, "tests", "golden", "idft.npz"))
    dev = torch.device("cuda")
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=dev)
    s, kxd, kyd = t(g["s"].reshape(1, -1), np.complex128), t(g["kx"], np.complex128), t(g["ky"], np.complex128)
    xd, yd = t(g["x"].ravel(), np.float64), t(g["y"].ravel(), np.float64)
    N, npts = s.shape[1], xd.numel()
    out = torch.empty((1, npts), dtype=torch.complex128, device=dev)
    ws = torch.empty(lib.kh_idft_work_bytes(N, npts), dtype=torch.uint8, device=dev)
    rc = lib.kh_idft_batch(1, N, npts, kxd.data_ptr(), kyd.data_ptr(), xd.data_ptr(), yd.data_ptr(), s.data_ptr(),
                           out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.kh_last_error()
    assert np.abs(out.cpu().numpy().reshape(g["x"].shape) - g["out"]).max() <= 1e-12
    # error behaviour: bad arguments and an undersized workspace are status codes, not crashes
    assert lib.kh_idft_batch(1, N, npts, None, kyd.data_ptr(), xd.data_ptr(), yd.data_ptr(), s.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), None) != 0
    assert lib.kh_idft_batch(1, N, npts, kxd.data_ptr(), kyd.data_ptr(), xd.data_ptr(), yd.data_ptr(), s.data_ptr(), out.data_ptr(), ws.data_ptr(), 16, None) != 0


def test_layer_api_solve_and_stack_layers():
    """The reference's Layer-level API (layer.py:35-60, 145-194; examples/layer_api): Layer.solve leaves S (and W, V, L, IC with
    `fields`) on the layer, stack_layers chains solved layers."""
    from khepri_b200 import Expansion, Layer
    from khepri_b200.layer import stack_layers
    from oracle import rcwa_oracle as orc
    from tests import cases
    from tests.util import engine
    eng = engine("emu")
    e = Expansion((3, 3))
    pm = cases.disc_pixmap((64, 64), 4.0, (0.0, 0.0), 0.25, 1.0)
    kp, wl = (0.3, -0.1), 1.3
    lp = Layer.pixmap(e, pm, 0.3)
    lp.fields = True
    lu = Layer.uniform(e, 2.2, 0.4)
    lp.solve(kp, wl, engine=eng)
    lu.solve(kp, wl, engine=eng)
    g = orc.g_vectors((3, 3), np.eye(2))
    rp = orc.solve_layer(("pixmap", pm, 0.3), g, (3, 3), kp, wl)
    ru = orc.solve_layer(("uniform", 2.2, 0.4), g, (3, 3), kp, wl)
    assert np.abs(lp.S - rp["S"]).max() < 1e-10 and np.abs(lu.S - ru["S"]).max() < 1e-12
    assert np.abs(lp.IC - rp["IC"]).max() < 1e-11 and lp.W.shape == (18, 18) and lp.L.shape == (18,)
    assert lu.W is None and lu.IC is None                                  # dropped without `fields` (layer.py:190-194)
    from tests.test_oracle_golden import match_spectrum
    match_spectrum(lp.L ** 2, rp["L"] ** 2, 1e-9)
    Sls, Srs, Stot = stack_layers((3, 3), [lp, lu, lp], [True, False, True], engine=eng)
    pre, suf, tot = orc.stack_chain(18, [rp["S"], ru["S"], rp["S"]])
    assert np.abs(Stot - tot).max() < 1e-10 and Sls[1] is None and np.abs(Sls[2] - pre[2]).max() < 1e-10
    assert np.abs(Srs[0] - suf[0]).max() < 1e-10
