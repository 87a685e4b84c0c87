"""CPU: pin the oracle (oracle/rcwa_oracle.py) against golden vectors produced by the unmodified
reference (oracle/gen_golden.py) and against the reference's own known-answer test."""
import os

import numpy as np
import pytest

from oracle import rcwa_oracle as orc
from tests import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def src_kp(st, s):
    if s.get("kp") is not None:
        return tuple(s["kp"])
    return tuple(orc.kplanar(st["epsi"], s["wavelength"], s.get("theta", 0.0), s.get("phi", 0.0)))


def oracle_sweep(st, srcs):
    out = []
    for s in srcs:
        sol = orc.solve_structure(st, s["wavelength"], src_kp(st, s), want_reverse=False)
        out.append(orc.flux_end(st, sol, s.get("te", 1.0), s.get("tm", 1.0)))
    return np.array(out)


def test_toeplitz_gather_bit_exact():
    g = gold("toeplitz")
    C = orc.toeplitz_gather(g["coded"], tuple(g["pw"]))
    assert np.array_equal(C, g["C"])


def test_convolution_matrix():
    g = gold("convmat")
    pm = cases.disc_pixmap((96, 64), 2.25, (0.05, -0.1), 0.3, 6.0)
    assert np.array_equal(orc.convolution_matrix(pm, (5, 3)), g["C"])
    pm = cases.disc_pixmap((128, 128), 12, (0, 0), 0.4, 1.0)
    assert np.array_equal(orc.convolution_matrix(pm, (7, 7)), g["C77"])


def test_suh03_spectrum():
    g = gold("suh03")
    st, srcs = cases.case_suh03()
    idx = list(range(0, 151, 6)) + [150]
    rt = oracle_sweep(st, [srcs[i] for i in idx])
    np.testing.assert_allclose(rt, g["RT"][idx], rtol=1e-10, atol=1e-12)
    for k, i in enumerate(g["Sidx"]):
        sol = orc.solve_structure(st, srcs[i]["wavelength"], src_kp(st, srcs[i]), want_reverse=False)
        np.testing.assert_allclose(sol["Stot"], g["Stot"][k], rtol=0, atol=1e-10)
        (_, rg), (_, tg) = orc.flux_end(st, sol, srcs[i]["te"], srcs[i]["tm"], only_total=False)
        np.testing.assert_allclose(np.stack([rg, tg]), g["orders"][i], rtol=0, atol=1e-11)


@pytest.mark.parametrize("pw,nk,nwl,tag", [((7, 7), 3, 3, "bzi77"), ((3, 3), 4, 5, "bzi33")])
def test_bzi(pw, nk, nwl, tag):
    st, srcs = cases.case_bzi(pw, nk, nwl)
    np.testing.assert_allclose(oracle_sweep(st, srcs), gold(tag)["RT"], rtol=1e-9, atol=1e-12)


def test_woodpile55_with_doubling():
    g = gold("woodpile55")
    st, srcs = cases.case_woodpile((5, 5), 3, 3)
    rt, rt2 = [], []
    for s in srcs:
        sol = orc.solve_structure(st, s["wavelength"], src_kp(st, s), want_reverse=False)
        rt.append(orc.flux_end(st, sol, s["te"], s["tm"]))
        sol["Stot"] = orc.star(sol["Stot"], sol["Stot"])
        rt2.append(orc.flux_end(st, sol, s["te"], s["tm"]))
    np.testing.assert_allclose(np.array(rt), g["RT"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.array(rt2), g["RT_doubled"], rtol=1e-9, atol=1e-12)


def test_woodpile1111_one_solve():
    g = gold("woodpile1111")
    st, srcs = cases.case_woodpile((11, 11), 2, 2)
    np.testing.assert_allclose(oracle_sweep(st, srcs[:1]), g["RT"][:1], rtol=1e-9, atol=1e-12)


def test_oblique_hexagonal_lossy():
    g = gold("oblique")
    st, srcs = cases.case_oblique()
    np.testing.assert_allclose(oracle_sweep(st, srcs), g["RT"], rtol=1e-10, atol=1e-12)


def test_fresnel_known_answer():
    g = gold("fresnel")
    fcases, rfres = cases.case_fresnel()
    rt = np.array([oracle_sweep(st, [src])[0] for st, src in fcases])
    np.testing.assert_allclose(rt, g["RT"], rtol=1e-12)
    np.testing.assert_allclose(rfres, rt[:, 0], rtol=1e-7)     # test_complex_eps.py:42


@pytest.mark.parametrize("pp,tag,tol", [(5, "fields55", 1e-9), (7, "fields77", 1e-8)])
def test_fields_volume(pp, tag, tol):
    g = gold(tag)
    st, src, (X, Y, z) = cases.case_fields(pp)
    sol = orc.solve_structure(st, src["wavelength"], src_kp(st, src))
    E, H = orc.fields_volume(st, sol, X, Y, z, src["te"], src["tm"])
    scale = np.abs(g["E"]).max()
    assert np.abs(E - g["E"]).max() <= tol * scale
    assert np.abs(H - g["H"]).max() <= tol * np.abs(g["H"]).max()
    np.testing.assert_allclose(orc.flux_end(st, sol, src["te"], src["tm"]), g["RT"], rtol=1e-10)


def test_fields_fourier_golden():
    """fields_coords_xy(..., return_fourier=True) (crystal.py:326-327) of the unmodified reference, normal and oblique source."""
    g = gold("fields55_fourier")
    st, src, (X, Y, z) = cases.case_fields(5)
    for s, key in ((src, "FF"), (dict(wavelength=1.9, te=0.6, tm=0.8, theta=17.0, phi=25.0), "FF_oblique")):
        kp = src_kp(st, s)
        sol = orc.solve_structure(st, s["wavelength"], kp)
        e = orc.incident_vector(st["pw"], s["te"], s["tm"], (kp[0], kp[1], orc.source_kzi(st, s["wavelength"], kp)))
        inc = np.concatenate([e, np.zeros_like(e)])
        FF = np.array([np.array(orc.fields_fourier_at(st, sol, zi, inc)) for zi in z])
        assert FF.shape == g[key].shape == (len(z), 6, 25)
        assert np.abs(FF - g[key]).max() <= 1e-9 * np.abs(g[key]).max()


def test_idft_golden():
    g = gold("idft")
    np.testing.assert_allclose(orc.idft(g["s"], g["kx"], g["ky"], g["x"], g["y"]), g["out"], rtol=0, atol=1e-13)


def test_twisted_bilayer_extended():
    g = gold("twisted33")
    tw = cases.twisted_case()
    g0 = orc.g_vectors(tw["pw"], np.eye(2))
    rt = np.empty((len(tw["freqs"]), len(tw["twists"]), 2))
    first = True
    for i, f in enumerate(tw["freqs"]):
        for j, ta in enumerate(tw["twists"]):
            desc = {"pw": tw["pw"], "g1": orc.rotation(ta / 2) @ g0, "g2": orc.rotation(-ta / 2) @ g0,
                    "epsi": 1, "epse": 1, "stack": ["upper", "inter", "lower"],
                    "layers": {"upper": (("pixmap", tw["pixmap"], tw["depths"][0]), 1),
                               "lower": (("pixmap", tw["pixmap"], tw["depths"][2]), 2),
                               "inter": (("uniform", 1, tw["depths"][1]), 1)}}
            sol = orc.solve_twisted(desc, 1 / f, (0j, 0j))
            if first:
                np.testing.assert_allclose(sol["layers"]["upper"]["S"][0, 0], g["S11_upper_first"], atol=1e-10)
                first = False
            st = {"pw": (tw["pw"][0] ** 2, tw["pw"][1] ** 2), "epsi": 1, "epse": 1}
            sol["layers"]["Sref"]["W"] = sol["layers"]["Strans"]["W"] = np.identity(sol["Stot"].shape[-1])
            rt[i, j] = orc.flux_end(st, sol, 1.0, 0.0)
    np.testing.assert_allclose(rt, g["RT"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("which", ["tidy", "mixed", "rect"])
def test_analytical_layers(which):
    """add_layer_analytical (analytic island transforms -> Toeplitz gather), layer.py:161-174."""
    g = gold("analytical")
    st, srcs = cases.case_analytical(which)
    name, spec = [(k, v) for k, v in st["layers"].items() if v[0] == "analytical"][0]
    C = orc.analytical_convolution_matrix(spec[1], spec[3], st["pw"], spec[4])
    assert np.abs(C - g["C_" + which]).max() <= 1e-14 * np.abs(g["C_" + which]).max()
    rt = oracle_sweep(st, srcs)
    assert np.abs(rt - g["RT_" + which]).max() <= 1e-10


def test_bzi_beam_amplitudes_and_summed_fields():
    """beams.amplitudes_from_fields + the k-sum of fields_volume (examples/bzi/bzi_animation.py:41-80)."""
    g = gold("bzi_beam")
    st, c = cases.case_bzi_beam()
    gv = orc.g_vectors(st["pw"], st["lattice"])
    xo, yo, zo = c["out"]
    total = 0
    for i, kp in enumerate(c["kbz"]):
        A = orc.beam_amplitudes(g["source"], gv, kp, c["X"], c["Y"], c["bz"])
        assert np.abs(A - g["amplitudes"][i]).max() <= 1e-12 * np.abs(g["amplitudes"]).max()
        sol = orc.solve_structure(st, c["wl"], tuple(kp))
        E, H = orc.fields_volume(st, sol, xo, yo, zo, None, None, incident_fields=A.reshape(-1))
        total = total + np.asarray((E, H))
    assert np.abs(total - g["fields"]).max() <= 1e-9 * np.abs(g["fields"]).max()


def match_spectrum(got, want, tol):
    """Every eigenvalue of `want` has a partner in `got` within tol relative (ordering is free, multiplicities respected)."""
    got = list(np.asarray(got))
    for w in np.asarray(want):
        d = [abs(g - w) for g in got]
        j = int(np.argmin(d))
        assert d[j] <= tol * max(1.0, abs(w)), (w, got[j], d[j])
        got.pop(j)


@pytest.mark.parametrize("tag", ["p5a", "p5b", "p3a", "p3b"])
def test_band_postprocessing(tag):
    """eigentricks.scattering_eigenvalues / scattering_det of the reference on Crystal-path S-matrices."""
    g = gold("bands")
    S = orc.flat_smatrix(g[tag + "_S"])
    w, v = orc.scattering_eigenvalues(S)
    match_spectrum(w, g[tag + "_w"], 1e-10)
    Sl, Sr = orc.scattering_splitlr(S)
    assert np.abs(Sl @ v - (Sr @ v) * w[None, :]).max() <= 1e-9 * np.abs(w).max()
    assert abs(orc.scattering_det(S) - g[tag + "_det"]) <= 1e-9 * abs(g[tag + "_det"])


def test_fields99_golden():
    """C5 at its own basis size (9x9, n = 162), volume on a small grid + one 256x256 plane on a strided subset."""
    g = gold("fields99")
    st, src, (X, Y, z), (XP, YP, zp, stride) = cases.case_fields_plane(9)
    sol = orc.solve_structure(st, src["wavelength"], src_kp(st, src))
    E, H = orc.fields_volume(st, sol, X, Y, z, src["te"], src["tm"])
    assert np.abs(E - g["E"]).max() <= 1e-8 * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= 1e-8 * np.abs(g["H"]).max()
    Ep, Hp = orc.fields_volume(st, sol, XP[::stride, ::stride], YP[::stride, ::stride], [zp], src["te"], src["tm"])
    assert np.abs(Ep[0] - g["Eplane"]).max() <= 1e-8 * np.abs(g["Eplane"]).max()
    assert np.abs(Hp[0] - g["Hplane"]).max() <= 1e-8 * np.abs(g["Hplane"]).max()
    np.testing.assert_allclose(orc.flux_end(st, sol, src["te"], src["tm"]), g["RT"], rtol=1e-9)


@pytest.mark.parametrize("pp", [13, 15])
def test_supercell_large_bases(pp):
    """C4 large set: direct 13x13 / 15x15 bases (n = 338 / 450), R, T and strided subsets of Stot[0,0], Stot[1,0]."""
    g = gold(f"supercell{pp}")
    st, srcs = cases.case_supercell(pp)
    for i, s in enumerate(srcs[:1 if pp == 15 else 2]):
        kp = src_kp(st, s)
        sol = orc.solve_structure(st, s["wavelength"], kp, want_reverse=False)
        np.testing.assert_allclose(orc.flux_end(st, sol, s["te"], s["tm"]), g["RT"][i], rtol=1e-9, atol=1e-12)
        assert np.abs(sol["Stot"][0, 0][::9, ::7] - g["S11"][i]).max() <= 1e-9 * np.abs(g["S11"][i]).max()
        assert np.abs(sol["Stot"][1, 0][::9, ::7] - g["S21"][i]).max() <= 1e-9 * np.abs(g["S21"][i]).max()


def test_legacy_expm_doubling_reproduces_the_crystal_path():
    """the reference's legacy layer algorithm (tmat/scattering.py:25-51: expm of a thin slice, matrix_s, self
    star products) restated in the Crystal basis reproduces the Crystal path's golden spectra without any eigensolver --
    the two algorithms of the reference agree, and either can check the CUDA path."""
    g = gold("suh03")
    st, srcs = cases.case_suh03()
    idx = [0, 37, 75, 112, 150]
    orc.PATTERNED_BY_DOUBLING = 4
    try:
        rt = oracle_sweep(st, [srcs[i] for i in idx])
        st2, srcs2 = cases.case_bzi((3, 3), 4, 5)
        rt2 = oracle_sweep(st2, srcs2[:6])
    finally:
        orc.PATTERNED_BY_DOUBLING = None
    np.testing.assert_allclose(rt, g["RT"][idx], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(rt2, gold("bzi33")["RT"][:6], rtol=1e-9, atol=1e-11)
