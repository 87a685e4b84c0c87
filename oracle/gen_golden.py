"""Generate golden vectors by running the UNMODIFIED reference (Kaeryv/Khepri).

Run in the build container only (the reference tree does not travel to the GPU box):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Writes small ``.npz`` fixtures to ``tests/golden/``.  Inputs come from ``tests/cases.py``
(deterministic, no RNG); pixmaps are rebuilt from the same helpers at test time, and this script
asserts that those helpers reproduce ``khepri.draw.Drawing`` bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from khepri.crystal import Crystal  # noqa: E402
from khepri.draw import Drawing  # noqa: E402
from khepri.expansion import Expansion  # noqa: E402
from khepri.layer import Layer  # noqa: E402
from khepri.tools import convolution_matrix, convolution_matrix_fourier  # noqa: E402
from khepri.alternative import redheffer_product  # noqa: E402
from khepri.misc import coords  # noqa: E402
from khepri.beams import gen_bzi_grid  # noqa: E402
from khepri.factory import make_woodpile  # noqa: E402

from tests import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_crystal(st, fields=False):
    cl = Crystal(st["pw"], lattice=st["lattice"], epsi=st["epsi"], epse=st["epse"])
    for name, spec in st["layers"].items():
        if spec[0] == "uniform":
            cl.add_layer_uniform(name, spec[1], spec[2])
        elif spec[0] == "analytical":
            cl.add_layer_analytical(name, spec[1], spec[3], spec[2])
        else:
            cl.add_layer_pixmap(name, spec[1], spec[2])
    cl.set_device(st["stack"], [fields] * len(st["stack"]))
    return cl


def sweep(st, srcs, per_order=False):
    cl = ref_crystal(st)
    rt, orders = [], []
    for s in srcs:
        cl.set_source(**s)
        cl.solve()
        rt.append(cl.poynting_flux_end())
        if per_order:
            (_, rg), (_, tg) = cl.poynting_flux_end(only_total=False)
            orders.append(np.stack([rg, tg]))
    return np.array(rt, dtype=float), (np.array(orders) if per_order else None), cl


def check_helpers():
    d = Drawing((128, 128), 12)
    d.disc((0, 0), 0.4, 1.0)
    assert np.array_equal(d.canvas(), cases.disc_pixmap((128, 128), 12, (0, 0), 0.4, 1.0))
    d = Drawing((128, 128), 1)
    d.rectangle((0, 0), (0.5, 1), 4)
    assert np.array_equal(d.canvas(), cases.rect_pixmap((128, 128), 1, (0, 0), (0.5, 1), 4))
    wp = make_woodpile(0.28, 3.6 ** 2, 0.5, 1.414 / 4, (3, 3), (256, 256))
    st = cases.woodpile_structure((3, 3))
    for k in "ABCD":
        assert np.array_equal(wp.layers[k].epsilon, st["layers"][k][1]), k
    assert np.array_equal(gen_bzi_grid((64, 64)), cases.bzi_kgrid((64, 64)))
    print("helpers reproduce Drawing / make_woodpile / gen_bzi_grid exactly")


def main():
    check_helpers()

    # --- convolution matrix: bit-exact gather on an index-coded array + FFT coefficients
    P, Q = 5, 3
    Nx, Ny = 16, 12
    coded = (np.arange(Nx)[:, None] * 1000 + np.arange(Ny)[None, :]).astype(complex) + 0.5j
    np.savez(os.path.join(OUT, "toeplitz.npz"), coded=coded, pw=np.array([P, Q]),
             C=convolution_matrix_fourier(coded, (P, Q)))
    pm = cases.disc_pixmap((96, 64), 2.25, (0.05, -0.1), 0.3, 6.0)
    np.savez(os.path.join(OUT, "convmat.npz"), pw=np.array([5, 3]), C=convolution_matrix(pm, (5, 3)),
             C77=convolution_matrix(cases.disc_pixmap((128, 128), 12, (0, 0), 0.4, 1.0), (7, 7)))
    print("convmat done")

    # --- C1: README suh03, full 151-frequency spectrum; plus Stot at three frequencies
    st, srcs = cases.case_suh03()
    rt, orders, cl = sweep(st, srcs, per_order=True)
    S_samples = []
    for i in (0, 75, 150):
        cl.set_source(**srcs[i])
        cl.solve()
        S_samples.append(np.asarray(cl.Stot))
    np.savez(os.path.join(OUT, "suh03.npz"), RT=rt, orders=orders, Stot=np.array(S_samples), Sidx=np.array([0, 75, 150]))
    print("suh03 done", rt[:2])

    # --- C2: BZI 7x7, sub-sampled k-grid x wavelengths
    st, srcs = cases.case_bzi((7, 7), nk=3, nwl=3)
    rt, _, _ = sweep(st, srcs)
    np.savez(os.path.join(OUT, "bzi77.npz"), RT=rt)
    st, srcs = cases.case_bzi((3, 3), nk=4, nwl=5)
    rt, _, _ = sweep(st, srcs)
    np.savez(os.path.join(OUT, "bzi33.npz"), RT=rt)
    print("bzi done")

    # --- C3: woodpile 11x11 (4 solves) and 5x5 (9 solves), incl. the Stot (x) Stot doubling (woodpile.py:85)
    for pw, nk, nf, tag in (((11, 11), 2, 2, "woodpile1111"), ((5, 5), 3, 3, "woodpile55")):
        st, srcs = cases.case_woodpile(pw, nk, nf)
        cl = ref_crystal(st)
        rt, rt2 = [], []
        for s in srcs:
            cl.set_source(**s)
            cl.solve()
            rt.append(cl.poynting_flux_end())
            cl.Stot = redheffer_product(cl.Stot, cl.Stot)
            rt2.append(cl.poynting_flux_end())
        np.savez(os.path.join(OUT, tag + ".npz"), RT=np.array(rt), RT_doubled=np.array(rt2))
    print("woodpile done")

    # --- oblique / hexagonal / lossy / epsi, epse != 1
    st, srcs = cases.case_oblique()
    rt, orders, _ = sweep(st, srcs, per_order=True)
    np.savez(os.path.join(OUT, "oblique.npz"), RT=rt, orders=orders)
    print("oblique done")

    # --- Fresnel (the reference's own asserted test)
    fcases, rfres = cases.case_fresnel()
    rt = []
    for st, src in fcases:
        cl = ref_crystal(st)
        cl.set_source(**src)
        cl.solve()
        rt.append(cl.poynting_flux_end())
    rt = np.array(rt)
    np.testing.assert_allclose(rfres, rt[:, 0])
    np.savez(os.path.join(OUT, "fresnel.npz"), RT=rt, R_fresnel=rfres)
    print("fresnel done")

    # --- C5: field maps on sliced holey pair (5x5 and 7x7), small xyz grid
    for pp, tag in ((5, "fields55"), (7, "fields77")):
        st, src, (X, Y, z) = cases.case_fields(pp)
        cl = ref_crystal(st, fields=True)
        cl.set_source(**src)
        cl.solve()
        E, H = cl.fields_volume(X, Y, z)
        np.savez(os.path.join(OUT, tag + ".npz"), E=E, H=H, RT=np.array(cl.poynting_flux_end()))
    x, y, z = coords(0, 1, 0, 1, 0.0001, 2.2, (12, 10, 9))
    st, src, (X, Y, Z) = cases.case_fields(5)
    assert np.array_equal(x, X) and np.array_equal(y, Y) and np.allclose(z, Z)
    print("fields done")

    # --- C4: twisted bilayer, extended RCWA (3,3)+(3,3)
    tw = cases.twisted_case()
    canvas = tw["pixmap"]
    d = Drawing((128, 128), 4)
    d.disc((0, 0), 0.25, 1)
    assert np.array_equal(d.canvas(), canvas)
    rt = []
    S_first = None
    for f in tw["freqs"]:
        for ta in tw["twists"]:
            e1, e2 = Expansion(tw["pw"]), Expansion(tw["pw"])
            e1.rotate(ta / 2)
            e2.rotate(-ta / 2)
            cl = Crystal.from_expansion(e1 + e2)
            cl.add_layer("upper_layer", Layer.pixmap(e1, canvas, tw["depths"][0]), extended=True)
            cl.add_layer("lower_layer", Layer.pixmap(e2, canvas, tw["depths"][2]), extended=True)
            cl.add_layer("interlayer", Layer.uniform(e1, 1, tw["depths"][1]), extended=True)
            cl.set_device(["upper_layer", "interlayer", "lower_layer"])
            cl.set_source(wavelength=1 / f, te=1, tm=0)
            cl.solve()
            rt.append(cl.poynting_flux_end())
            if S_first is None:
                S_first = np.asarray(cl.layers["upper_layer"].S)[0, 0]
    np.savez(os.path.join(OUT, "twisted33.npz"), RT=np.array(rt).reshape(len(tw["freqs"]), len(tw["twists"]), 2),
             S11_upper_first=S_first)
    print("twisted done")


def gen_analytical():
    """layers from analytic island transforms (Crystal.add_layer_analytical, layer.py:161-174)."""
    d = Drawing((64, 64), 2.2)
    d.rectangle((0.1, -0.05), (0.5, 0.3), 6.0)
    d.disc((-0.2, 0.15), 0.12, 1.0)
    mine = [cases.rect_island((0.1, -0.05), (0.5, 0.3), 6.0), cases.disc_island((-0.2, 0.15), 0.12, 1.0)]
    for a, b in zip(d.islands(), mine):
        assert a["type"] == b["type"] and np.array_equal(np.asarray(a["params"], float), np.asarray(b["params"], float)) and a["epsilon"] == b["epsilon"]
    out = {}
    for which in ("tidy", "mixed", "rect"):
        st, srcs = cases.case_analytical(which)
        rt, _, cl = sweep(st, srcs)
        name = [k for k, v in st["layers"].items() if v[0] == "analytical"][0]
        out["RT_" + which] = rt
        out["C_" + which] = np.asarray(cl.layers[name].C)
    np.savez(os.path.join(OUT, "analytical.npz"), **out)
    print("analytical done", out["RT_tidy"][:2])


def gen_bzi_beam():
    """BZI source amplitudes (beams.amplitudes_from_fields) and the k-summed field maps (bzi_animation.py:41-80)."""
    from khepri.beams import _paraxial_gaussian_field_fn, shifted_rotated_fields, amplitudes_from_fields
    st, c = cases.case_bzi_beam()
    b = c["beam"]
    X, Y = c["X"], c["Y"]
    src = shifted_rotated_fields(_paraxial_gaussian_field_fn, X, Y, np.zeros_like(X), b["wl"], b["x0"], b["y0"], b["z0"],
                                 b["theta"], b["phi"], b["pol"], beam_waist=b["beam_waist"], er=b["er"])
    src = np.asarray(src)
    if src.ndim == 5:                                                            # numpy >= 2: solve() keeps a trailing singleton (beams.py:69-71)
        src = src[..., 0]
    src = np.swapaxes(np.swapaxes(src, 0, 2), 1, 3)                             # (ny, nx, 2, 3) as in the example
    e1 = Expansion(st["pw"])
    xo, yo, zo = c["out"]
    amps, total = [], None
    for kp in c["kbz"]:
        cl = ref_crystal(st, fields=True)
        cl.set_source(c["wl"], np.nan, np.nan, kp=tuple(kp))
        cl.solve()
        F = amplitudes_from_fields(src, e1, c["wl"], tuple(kp), X, Y, c["bz"])
        amps.append(F)
        S, U = np.split(F.flatten(), 2)
        E, H = cl.fields_volume(xo, yo, zo, incident_fields=(S, U))
        total = np.asarray((E, H)) if total is None else total + np.asarray((E, H))
    np.savez(os.path.join(OUT, "bzi_beam.npz"), source=src, amplitudes=np.array(amps), fields=total)
    print("bzi beam done", np.abs(total).max())



def gen_bands():
    """Band post-processing: khepri.eigentricks on unit-cell S-matrices of the Crystal path."""
    from khepri.eigentricks import scattering_eigenvalues as ref_eigs
    out = {}
    st5, srcs5 = cases.case_suh03()
    st3 = cases.holey_pair(3, 128)
    todo = [("p5a", st5, srcs5[10]), ("p5b", st5, dict(srcs5[120], theta=20.0, phi=30.0)),
            ("p3a", st3, dict(wavelength=1 / 0.52, te=1.0, tm=0.0)), ("p3b", st3, dict(wavelength=1 / 0.58, te=1.0, tm=1.0, theta=35.0, phi=10.0))]
    for tag, st, src in todo:
        cl = ref_crystal(st)
        cl.set_source(**src)
        cl.solve()
        S4 = np.asarray(cl.Stot)
        S = np.block([[S4[0, 0], S4[0, 1]], [S4[1, 0], S4[1, 1]]])
        w, v, det = ref_eigs(S, dos=True)
        out[tag + "_S"] = S4
        out[tag + "_w"] = w
        out[tag + "_det"] = np.asarray(det)
    np.savez_compressed(os.path.join(OUT, "bands.npz"), **out)
    print("bands.npz", {k: v.shape for k, v in out.items()})


def gen_fields_fourier():
    """fields_coords_xy(..., return_fourier=True) (crystal.py:326-327): the six Fourier vectors at every depth of the
    fields55 case, plus one oblique two-polarisation source with explicit incident fields."""
    st, src, (X, Y, z) = cases.case_fields(5)
    cl = ref_crystal(st, fields=True)
    cl.set_source(**src)
    cl.solve()
    FF = np.array([np.array(cl.fields_coords_xy(X, Y, zi, return_fourier=True)) for zi in z])
    src2 = dict(wavelength=1.9, te=0.6, tm=0.8, theta=17.0, phi=25.0)
    cl.set_source(**src2)
    cl.solve()
    FF2 = np.array([np.array(cl.fields_coords_xy(X, Y, zi, return_fourier=True)) for zi in z])
    np.savez(os.path.join(OUT, "fields55_fourier.npz"), FF=FF, FF_oblique=FF2, z=np.asarray(z))
    # fourier.idft itself on scattered points with complex k (deterministic inputs, no RNG state shared with the tests)
    from khepri.fourier import idft
    rng = np.random.default_rng(11)
    sv = rng.standard_normal(25) + 1j * rng.standard_normal(25)
    kx = rng.standard_normal(25) * 4 + 0.05j * rng.standard_normal(25)
    ky = rng.standard_normal(25) * 4 + 0j
    xx, yy = rng.random((6, 7)), rng.random((6, 7)) - 0.5
    np.savez(os.path.join(OUT, "idft.npz"), s=sv, kx=kx, ky=ky, x=xx, y=yy, out=idft(sv, kx, ky, xx, yy))
    print("fields fourier done", FF.shape)


def gen_round2():
    """Round 2: the two BASELINE configs that had no golden at their own size.
    C5 at 9x9 (n = 162: blocked inverse / tiled Hessenberg inside the field pipeline), volume on a small grid plus one
    256 x 256 plane kept on a strided subset; C4 large set: direct 13x13 / 15x15 bases (n = 338 / 450), R, T and a strided
    subset of Stot[0,0] / Stot[1,0] for two sources each."""
    import time
    st, src, (X, Y, z), (XP, YP, zp, stride) = cases.case_fields_plane(9)
    cl = ref_crystal(st, fields=True)
    cl.set_source(**src)
    t0 = time.time()
    cl.solve()
    E, H = cl.fields_volume(X, Y, z)
    Ep, Hp = cl.fields_coords_xy(XP, YP, zp)
    np.savez(os.path.join(OUT, "fields99.npz"), E=E, H=H, RT=np.array(cl.poynting_flux_end()),
             Eplane=np.asarray(Ep)[:, ::stride, ::stride], Hplane=np.asarray(Hp)[:, ::stride, ::stride], zplane=zp, stride=stride)
    print("fields99 done", time.time() - t0, np.abs(E).max())
    for pp in (13, 15):
        st, srcs = cases.case_supercell(pp)
        cl = ref_crystal(st)
        rt, s11, s21 = [], [], []
        for sc in srcs:
            t0 = time.time()
            cl.set_source(**sc)
            cl.solve()
            rt.append(cl.poynting_flux_end())
            S = np.asarray(cl.Stot)
            s11.append(S[0, 0][::9, ::7].copy())
            s21.append(S[1, 0][::9, ::7].copy())
            print(pp, "solve", time.time() - t0, rt[-1])
        np.savez(os.path.join(OUT, f"supercell{pp}.npz"), RT=np.array(rt), S11=np.array(s11), S21=np.array(s21))


def gen_twisted_fields():
    """Field maps of a twisted bilayer (extended layers with retained eigenspaces, extension.py:100-112; crystal.py:250-253)."""
    tw = cases.twisted_case()
    out = {}
    for tag, it, jf, src in (("a", 1, 1, dict(te=1, tm=0)), ("b", 2, 0, dict(te=0.5, tm=1.0, theta=12.0, phi=20.0))):
        e1, e2 = Expansion(tw["pw"]), Expansion(tw["pw"])
        ta = tw["twists"][it]
        e1.rotate(ta / 2)
        e2.rotate(-ta / 2)
        cl = Crystal.from_expansion(e1 + e2)
        cl.add_layer("upper_layer", Layer.pixmap(e1, tw["pixmap"], tw["depths"][0]), extended=True)
        cl.add_layer("lower_layer", Layer.pixmap(e2, tw["pixmap"], tw["depths"][2]), extended=True)
        cl.add_layer("interlayer", Layer.uniform(e1, 1, tw["depths"][1]), extended=True)
        cl.set_device(["upper_layer", "interlayer", "lower_layer"], [True] * 3)
        cl.set_source(wavelength=1 / tw["freqs"][jf], **src)
        cl.solve()
        X, Y, z = cases.twisted_field_grid()
        E, H = cl.fields_volume(X, Y, z)
        out["E_" + tag], out["H_" + tag], out["RT_" + tag] = E, H, np.array(cl.poynting_flux_end())
    np.savez(os.path.join(OUT, "twisted33_fields.npz"), **out)
    print("twisted fields done", np.abs(out["E_a"]).max())


if __name__ == "__main__":
    if "--twisted-fields" in sys.argv:
        gen_twisted_fields()
    elif "--round2" in sys.argv:
        gen_round2()
    elif "--fields-fourier" in sys.argv:
        gen_fields_fourier()
    elif "--bzi-beam" in sys.argv:
        gen_bzi_beam()
    elif "--analytical" in sys.argv:
        gen_analytical()
    elif "--bands" in sys.argv:
        gen_bands()
    else:
        main()
        gen_analytical()
        gen_bzi_beam()
        gen_bands()
        gen_fields_fourier()
        gen_round2()
        gen_twisted_fields()
