"""Deterministic synthetic structures and sources of the BASELINE.json configs: input builders for
bench.py, the golden generator, the oracle tests and the GPU parity tests.  Pure numpy, no compute.

A *structure* is the plain dict of oracle/rcwa_oracle.py; a *case* = (structure, list of sources), each source a dict of
``set_source`` keyword arguments (khepri/crystal.py:345-360).  Test cases are sub-sampled so the oracle finishes in seconds.
"""
import numpy as np


def disc_pixmap(shape, eps_bg, center, radius, eps):
    xs = np.linspace(-0.5, 0.5, shape[0])[:, None]
    ys = np.linspace(-0.5, 0.5, shape[1])[None, :]
    pm = np.full(shape, float(eps_bg))
    pm[np.sqrt((xs - center[0]) ** 2 + (ys - center[1]) ** 2) < radius] = eps
    return pm


def rect_pixmap(shape, eps_bg, center, wh, eps, base=None):
    xs = np.linspace(-0.5, 0.5, shape[0])[:, None]
    ys = np.linspace(-0.5, 0.5, shape[1])[None, :]
    pm = np.full(shape, float(eps_bg)) if base is None else base
    x0, y0 = center[0] - wh[0] / 2, center[1] - wh[1] / 2
    m = (xs >= x0) & (xs <= x0 + wh[0]) & (ys >= y0) & (ys <= y0 + wh[1])
    pm[m] = eps
    return pm


def _st(pw, layers, stack, lattice=None, epsi=1, epse=1):
    lattice = np.eye(2) if lattice is None else np.asarray(lattice, dtype=float)
    return {"pw": tuple(pw), "lattice": lattice, "epsi": epsi, "epse": epse,
            "layers": dict(layers), "stack": list(stack)}


def holey_pair(pp=5, res=128, slices=1):
    """README suh03 / examples/crystal_api/test_crystal.py:23-32 (C1, C5)."""
    pm = disc_pixmap((res, res), 12, (0.0, 0.0), 0.4, 1.0)
    layers = {"S1": ("uniform", 1, 1.1 / slices), "Scyl": ("pixmap", pm, 0.55 / slices)}
    stack = ["Scyl"] * slices + ["S1"] * slices + ["Scyl"] * slices
    return _st((pp, pp), layers, stack)


def case_suh03(nf=151):
    st = holey_pair(5, 128)
    freqs = np.linspace(0.49, 0.6, 151)
    if nf < 151:
        freqs = freqs[:: max(1, 151 // nf)][:nf]
    return st, [dict(wavelength=1 / f, te=1.0, tm=0.0, theta=0.0, phi=0.0) for f in freqs]


def bzi_structure(pw=(7, 7)):
    """examples/bzi/bzi_animation.py:55-68 (C2)."""
    pm = rect_pixmap((128, 128), 1, (0, 0), (0.5, 1), 4)
    layers = {"S1": ("uniform", 1, 0.99), "S2": ("uniform", 4, 16.99), "S3": ("pixmap", pm, 0.4)}
    return _st(pw, layers, ["S1"] * 14 + ["S3", "S2"], epsi=1, epse=4)


def bzi_kgrid(shape):
    """khepri/beams.py:193-207 (square lattice, a=1)."""
    si, sj = 1 / shape[0], 1 / shape[1]
    i, j = np.meshgrid(np.arange(-0.5 + si / 2, 0.5, si), np.arange(-0.5 + sj / 2, 0.5, sj), indexing="ij")
    return np.stack([2 * np.pi * i, 2 * np.pi * j])


def case_bzi(pw=(7, 7), nk=2, nwl=3):
    st = bzi_structure(pw)
    kg = bzi_kgrid((64, 64))
    wls = 1 / np.linspace(0.8, 1.0, 101)
    srcs = []
    for a in np.linspace(3, 60, nk).astype(int):
        for w in wls[:: max(1, 101 // nwl)][:nwl]:
            srcs.append(dict(wavelength=float(w), te=1.0, tm=1.0, kp=(float(kg[0, a, (a * 7) % 64]), float(kg[1, a, (a * 7) % 64]))))
    return st, srcs


def woodpile_structure(pw=(11, 11), res=(256, 256)):
    """khepri/factory.py:3-24 with examples/crystal_api/woodpile.py:36-39 parameters (C3)."""
    w, eps, shift, h = 0.28, 3.6 ** 2, 0.5, 1.414 / 4
    p1 = rect_pixmap(res, 1, (0, 0), (1, w), eps)
    p2 = rect_pixmap(res, 1, (0, shift), (1, w), eps)
    p2 = rect_pixmap(res, 1, (0, -shift), (1, w), eps, base=p2)
    layers = {"A": ("pixmap", p1, h), "B": ("pixmap", p1.T.copy(), h),
              "C": ("pixmap", p2, h), "D": ("pixmap", p2.T.copy(), h)}
    return _st(pw, layers, ["A", "B", "C", "D"])


def case_woodpile(pw=(11, 11), nk=2, nf=2):
    st = woodpile_structure(pw)
    freqs = np.linspace(0.4 / 1.414, 0.65 / 1.414, 200)
    kxs = np.linspace(0, 0.99 * np.pi, 200)
    srcs = []
    for ik in np.linspace(5, 100, nk).astype(int):        # below the light line (flux defined)
        for jf in np.linspace(10, 180, nf).astype(int):
            srcs.append(dict(wavelength=float(1 / freqs[jf]), te=1.0, tm=1.0, kp=(float(kxs[ik]), 0.0)))
    return st, srcs


def case_oblique():
    """Hexagonal lattice, lossy uniform layer, epsi/epse != 1, oblique incidence, both polarisations."""
    lat = 0.9 * np.array([[np.sqrt(3) / 2, 0.5], [np.sqrt(3) / 2, -0.5]])
    pm = disc_pixmap((96, 64), 2.25, (0.05, -0.1), 0.3, 6.0)
    layers = {"U": ("uniform", 2.1 - 0.3j, 0.37), "G": ("pixmap", pm, 0.21), "V": ("uniform", 1.7, 0.15)}
    st = _st((5, 3), layers, ["U", "G", "V", "G"], lattice=lat, epsi=1.44, epse=2.25)
    srcs = [dict(wavelength=wl, te=te, tm=tm, theta=th, phi=ph)
            for wl, te, tm, th, ph in [(1.31, 1.0, 0.0, 12.0, 0.0), (1.31, 0.0, 1.0, 12.0, 30.0),
                                       (0.93, 0.7, 0.4, 35.0, 75.0), (1.77, 1.0, 1.0, 5.0, -20.0),
                                       (0.81, 0.3, 1.0, 50.0, 10.0)]]
    return st, srcs


def case_fresnel():
    """test/integration/test_complex_eps.py:14-42 (pw=(1,1) lossy slab; closed-form Fresnel check)."""
    covera = 299792458 / 1e-6
    e0, sigma, h = 8.85418782e-12, 0.01e6, 1.2
    wls = np.linspace(0.7, 2.0)
    cases = []
    for wl in wls:
        omega = 2 * np.pi * covera / wl
        eps = 1.6 ** 2 - 1j * sigma / omega / e0
        cases.append((_st((1, 1), {"1": ("uniform", eps, h)}, ["1"]), dict(wavelength=float(wl), te=1, tm=1)))
    omega = 2 * np.pi * covera / wls
    eps = 1.6 ** 2 + 1j * sigma / omega / e0
    n2 = np.conj(np.sqrt(eps))
    r12, r23 = (1 - n2) / (1 + n2), (n2 - 1) / (n2 + 1)
    ph = np.exp(-2j * 2 * np.pi / wls * n2 * h)
    return cases, np.abs((r12 + r23 * ph) / (1 + r12 * r23 * ph)) ** 2


def case_fields(pp=5, slices=4, res=128, grid=(12, 10, 9)):
    """C5 geometry (holey pair, sliced for conditioning -- SURVEY.md §7.5) on a small xyz grid."""
    st = holey_pair(pp, res, slices)
    x = np.linspace(0, 1, grid[0])
    y = np.linspace(0, 1, grid[1])
    depth = 0.55 * 2 + 1.1
    z = np.linspace(0.0001, depth, grid[2])
    X, Y = np.meshgrid(x, y, indexing="xy")
    src = dict(wavelength=1 / 0.53, te=1.0, tm=0.0, theta=0.0, phi=0.0)
    return st, src, (X, Y, z)


def two_layer_structure(pp, res=256):
    """Direct supercell basis with two different pixmap layers around a uniform spacer (the large end of C4:
    notebooks/Twisted.ipynb cells 6-8 solve a (15, 15) Crystal with two 512^2 pixmap layers of depth 0.2)."""
    pm = disc_pixmap((res, res), 4.0, (0.0, 0.0), 0.25, 1.0)
    pm2 = pm.T.copy() + 0.5 * rect_pixmap((res, res), 0.0, (0.1, 0.0), (0.3, 0.5), 1.0)
    layers = {"A": ("pixmap", pm, 0.2), "B": ("pixmap", pm2, 0.2), "U": ("uniform", 1.0, 0.3)}
    return _st((pp, pp), layers, ["A", "U", "B"])


def case_supercell(pp):
    """Two sources on the direct pp x pp basis (n = 2 pp^2: 338 at 13x13, 450 at 15x15): normal and oblique incidence."""
    st = two_layer_structure(pp)
    srcs = [dict(wavelength=1 / 0.74, te=1.0, tm=0.0, theta=0.0, phi=0.0), dict(wavelength=1 / 0.81, te=0.6, tm=0.8, theta=14.0, phi=25.0)]
    return st, srcs


def case_fields_plane(pp=9, npl=256, stride=8):
    """C5 at its own basis size: 9x9 harmonics, sliced holey pair, plus one npl x npl plane (compared on a strided subset)."""
    st, src, (X, Y, z) = case_fields(pp, slices=4, res=128)
    xp = np.linspace(0, 1, npl)
    XP, YP = np.meshgrid(xp, xp, indexing="xy")
    return st, src, (X, Y, z), (XP, YP, 0.8, stride)


def twisted_case(pw=(3, 3), nf=3, nt=3):
    """notebooks/PRL_2021_BL.ipynb cells 2-8 (C4 parity set), sub-sampled."""
    pm = disc_pixmap((128, 128), 4, (0, 0), 0.25, 1.0)
    freqs = np.linspace(0.7, 0.83, 50)[:: 50 // nf][:nf]
    twists = np.deg2rad(np.linspace(0, 45, 50))[3:: 50 // nt][:nt]
    return {"pw": pw, "pixmap": pm, "depths": (0.2, 0.3, 0.2), "freqs": freqs, "twists": twists}


def twisted_field_grid():
    """Small xyz grid through the three layers of the twisted bilayer (depths 0.2 / 0.3 / 0.2)."""
    X, Y = np.meshgrid(np.linspace(0, 1, 6), np.linspace(0, 1, 5), indexing="xy")
    return X, Y, np.linspace(0.01, 0.69, 7)


def rect_island(center, wh, eps):
    """khepri/draw.py:46-55 (Drawing.rectangle's geometric description)."""
    x, y = center[0] - wh[0] / 2, center[1] - wh[1] / 2
    return {"type": "rectangle", "params": [0.5 + x, 0.5 + y, 0.5 + x + wh[0], 0.5 + y + wh[1]], "epsilon": eps}


def disc_island(center, radius, eps):
    """khepri/draw.py:41-44."""
    return {"type": "disc", "params": [0.5 + center[0], 0.5 + center[1], radius], "epsilon": eps}


def case_analytical(which="tidy"):
    """Layers from analytic island transforms (Crystal.add_layer_analytical).
    tidy : test/integration/test_tidy.py:13-31 (square rod eps 4 in air, depth 1, 13 wavelengths, pol (1,1)), 7x7 harmonics
    mixed: asymmetric rectangle + disc in a host of eps 2.2 over a uniform slab, oblique incidence, 5x5
    rect : pw = (3, 5), exposes the reference's reshape of the coefficient table for P != Q"""
    lat = np.eye(2)
    if which == "tidy":
        layers = {"1": ("analytical", [rect_island((0, 0), (0.5, 0.5), 4)], 1.0, 1, lat)}
        st = _st((7, 7), layers, ["1"])
        srcs = [dict(wavelength=float(w), te=1.0, tm=1.0, theta=0.0, phi=0.0) for w in np.linspace(1.01, 2, 13)]
    elif which == "mixed":
        isl = [rect_island((0.1, -0.05), (0.5, 0.3), 6.0), disc_island((-0.2, 0.15), 0.12, 1.0)]
        layers = {"A": ("analytical", isl, 0.4, 2.2, lat), "U": ("uniform", 1.5, 0.3)}
        st = _st((5, 5), layers, ["A", "U", "A"], epse=2.0)
        srcs = [dict(wavelength=float(w), te=0.7, tm=0.4, theta=12.0, phi=33.0) for w in np.linspace(1.2, 1.9, 9)]
    else:
        isl = [rect_island((0.05, 0.1), (0.4, 0.6), 5.0)]
        layers = {"A": ("analytical", isl, 0.5, 1.0, lat)}
        st = _st((3, 5), layers, ["A"])
        srcs = [dict(wavelength=float(w), te=1.0, tm=0.5, theta=5.0, phi=10.0) for w in np.linspace(1.3, 1.8, 6)]
    return st, srcs


def case_bzi_beam(bz=(5, 1), NS=7, pw=(3, 1)):
    """examples/bzi/bzi_animation.py at test size: 1-D grating (pw = (P, 1)), Gaussian beam at 25 degrees sampled on a
    supercell of bz unit cells with NS x NS samples each, BZ grid of bz k-points shifted by the beam's k-parallel."""
    wl, theta, eps1, eps2, zmax = 1.1, np.deg2rad(25.0), 1.0, 4.0, 6.0
    pm = rect_pixmap((64, 64), eps1, (0, 0), (0.5, 1), eps2)
    layers = {"S1": ("uniform", eps1, 0.99), "S3": ("pixmap", pm, 0.4), "S2": ("uniform", eps2, 2.99)}
    st = _st(pw, layers, ["S1"] * 2 + ["S3", "S2"], epsi=eps1, epse=eps2)
    X, Y = np.meshgrid(np.linspace(0, bz[0], NS * bz[0], endpoint=True), np.linspace(0, bz[1], NS * bz[1], endpoint=True))
    si, sj = 1 / bz[0], 1 / bz[1]
    i, j = np.meshgrid(np.arange(-0.5 + si / 2, 0.5, si), np.arange(-0.5 + sj / 2, 0.5, sj), indexing="ij")
    kbz = np.stack([2 * np.pi * i, 2 * np.pi * j]).reshape(2, -1).T.copy()
    kbz[:, 0] += np.sqrt(eps1) * 2 * np.pi / wl * np.sin(theta)
    xo, yo = np.meshgrid(np.linspace(0, bz[0], 24), np.linspace(bz[1] / 2, bz[1] / 2, 3), indexing="xy")
    zo = np.linspace(0.01, zmax, 5)
    beam = dict(wl=wl, x0=bz[0] / 2, y0=bz[1] / 2, z0=-zmax / 2, theta=theta, phi=0.0, pol=np.pi / 2, beam_waist=2 * wl, er=eps1)
    return st, dict(wl=wl, bz=bz, NS=NS, X=X, Y=Y, kbz=kbz, beam=beam, out=(xo, yo, zo))


def build_crystal(st, engine=None, fields=False, method="auto", crystal_cls=None):
    """A Crystal (khepri_b200's by default; pass the reference's class for the CPU arm) from a structure dict."""
    if crystal_cls is None:
        from khepri_b200 import Crystal as crystal_cls
        cl = crystal_cls(st["pw"], lattice=st["lattice"], epsi=st["epsi"], epse=st["epse"], engine=engine)
        cl.method = method
    else:
        cl = crystal_cls(st["pw"], lattice=st["lattice"], epsi=st["epsi"], epse=st["epse"])
    for name, spec in st["layers"].items():
        if spec[0] == "uniform":
            cl.add_layer_uniform(name, spec[1], spec[2])
        elif spec[0] == "analytical":
            cl.add_layer_analytical(name, spec[1], spec[3], spec[2])
        else:
            cl.add_layer_pixmap(name, spec[1], spec[2])
    cl.set_device(st["stack"], [fields] * len(st["stack"]))
    return cl
