"""Field maps (configs[4]), extended twisted-bilayer RCWA (configs[3]) and multi-rank sharding."""
import os
import sys

import numpy as np
import pytest

from oracle import rcwa_oracle as orc
from tests import cases
from tests.util import BACKENDS, ROOT, build_crystal, engine, gold


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("pp,tag,tol", [(5, "fields55", 1e-9), (7, "fields77", 1e-8)])
def test_fields_volume_golden(backend, pp, tag, tol):
    """fields_volume on the sliced holey pair vs the unmodified reference (1e-9 of max|E|; the 7x7
    reference itself is only reproducible to ~1e-8 on this stack)."""
    if backend == "emu" and pp == 7:
        pytest.skip("covered on the GPU; slow in emulation")
    eng = engine(backend)
    g = gold(tag)
    st, src, (X, Y, z) = cases.case_fields(pp)
    cl = build_crystal(st, eng, fields=True)
    cl.set_source(**src)
    cl.solve()
    E, H = cl.fields_volume(X, Y, z)
    assert E.shape == g["E"].shape == (len(z), 3) + X.shape
    assert np.abs(E - g["E"]).max() <= tol * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= tol * np.abs(g["H"]).max()
    Exy, Hxy = cl.fields_coords_xy(X, Y, z[4])
    assert np.abs(Exy - g["E"][4]).max() <= tol * np.abs(g["E"]).max()
    np.testing.assert_allclose(cl.poynting_flux_end(), g["RT"], rtol=1e-9)
    # stored partial products follow the reference's layout (layer.py:35-60)
    assert len(cl.stacking_matrices) == len(cl.global_stacking) == len(cl.stacking_reverse_matrices)
    assert np.abs(cl.stacking_matrices[-1] - cl.Stot).max() < 1e-12
    assert np.abs(cl.stacking_reverse_matrices[0][0, 0] - orc.star(orc.identity_smatrix(2 * pp * pp), cl.stacking_reverse_matrices[0])[0, 0]).max() < 1e-12


@pytest.mark.gpu
def test_fields_volume_9x9_golden():
    """C5 at its own basis size (9x9 harmonics, n = 162: blocked inverse and tiled Hessenberg inside the field pipeline):
    volume on a small grid and one 256 x 256 plane (separable grid kernel), compared with the unmodified reference on a
    strided subset of the plane.  Sliced stack, so the reference itself is reproducible to ~1e-9."""
    eng = engine("cuda")
    g = gold("fields99")
    st, src, (X, Y, z), (XP, YP, zp, stride) = cases.case_fields_plane(9)
    cl = build_crystal(st, eng, fields=True)
    cl.set_source(**src)
    cl.solve()
    E, H = cl.fields_volume(X, Y, z)
    assert np.abs(E - g["E"]).max() <= 1e-9 * np.abs(g["E"]).max()
    assert np.abs(H - g["H"]).max() <= 1e-9 * np.abs(g["H"]).max()
    Ep, Hp = cl.fields_coords_xy(XP, YP, zp)
    assert Ep.shape == (3, 256, 256)
    assert np.abs(Ep[:, ::stride, ::stride] - g["Eplane"]).max() <= 1e-9 * np.abs(g["Eplane"]).max()
    assert np.abs(Hp[:, ::stride, ::stride] - g["Hplane"]).max() <= 1e-9 * np.abs(g["Hplane"]).max()
    np.testing.assert_allclose(cl.poynting_flux_end(), g["RT"], rtol=1e-9)


@pytest.mark.parametrize("backend", BACKENDS)
def test_fields_return_fourier_golden(backend):
    """fields_coords_xy(..., return_fourier=True) -> (sx, sy, sz, ux, uy, uz) vs the unmodified reference (crystal.py:326-327),
    and consistency with the real-space maps through the oracle's idft (fourier.py:136-142)."""
    eng = engine(backend)
    g = gold("fields55_fourier")
    st, src, (X, Y, z) = cases.case_fields(5)
    cl = build_crystal(st, eng, fields=True)
    for s, key in ((src, "FF"), (dict(wavelength=1.9, te=0.6, tm=0.8, theta=17.0, phi=25.0), "FF_oblique")):
        cl.set_source(**s)
        cl.solve()
        for iz in (0, 4, len(z) - 1):
            ff = cl.fields_coords_xy(X, Y, z[iz], return_fourier=True)
            assert isinstance(ff, tuple) and len(ff) == 6 and ff[0].shape == (25,)
            assert np.abs(np.array(ff) - g[key][iz]).max() <= 1e-9 * np.abs(g[key]).max()
        E, _ = cl.fields_coords_xy(X, Y, z[4])
        k0 = 2 * np.pi / s["wavelength"]
        Kx, Ky, _ = cl.expansion.k_vectors(cl.kp, s["wavelength"])
        ff = cl.fields_coords_xy(X, Y, z[4], return_fourier=True)
        Ex = orc.idft(ff[0], k0 * Kx, k0 * Ky, X, Y)
        assert np.abs(Ex - E[0]).max() <= 1e-10 * max(1.0, np.abs(E).max())
    with pytest.raises(NotImplementedError):
        cl.fields_coords_xy(X, Y, "farfield")


@pytest.mark.parametrize("backend", BACKENDS)
def test_idft_operator(backend):
    """khepri_b200.fourier.idft == fourier.py:136-142 (oracle restatement) on scattered points, complex k, stacks and ragged sizes."""
    from khepri_b200.fourier import idft
    eng = engine(backend)
    g = gold("idft")                                                     # output of the unmodified reference
    assert np.abs(idft(g["s"], g["kx"], g["ky"], g["x"], g["y"], engine=eng) - g["out"]).max() <= 1e-12
    rng = np.random.default_rng(7)
    for N, shape, M in ((25, (6, 7), 1), (9, (1, 1), 3), (49, (5, 13), 6), (1, (3, 2), 2)):
        s = rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N))
        kx = rng.standard_normal(N) * 4 + 0.05j * rng.standard_normal(N)
        ky = rng.standard_normal(N) * 4 + 0j
        x, y = rng.random(shape), rng.random(shape) - 0.5
        want = np.array([orc.idft(s[m], kx, ky, x, y) for m in range(M)])
        got = idft(s, kx, ky, x, y, engine=eng)
        assert got.shape == (M,) + shape
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        one = idft(s[0], kx, ky, x, y, engine=eng)                       # the reference's call shape
        assert one.shape == shape and np.abs(one - want[0]).max() <= 1e-12 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("backend", BACKENDS)
def test_fields_need_retained_eigenspace(backend):
    eng = engine(backend)
    st, src, (X, Y, z) = cases.case_fields(5, slices=1)
    cl = build_crystal(st, eng, fields=False)
    cl.set_source(**src)
    cl.solve()
    with pytest.raises(AssertionError):
        cl.fields_volume(X, Y, z)
    with pytest.raises(AssertionError):
        cl.fields_coords_xy(X[0], Y[0], 0.1)          # x and y must be 2D meshgrids


@pytest.mark.parametrize("backend", BACKENDS)
def test_fields_oracle_explicit_incident_and_oblique(backend):
    """Oblique incidence + both polarisations, checked against the oracle (no golden needed)."""
    eng = engine(backend)
    st, _, (X, Y, z) = cases.case_fields(5, slices=2, grid=(7, 6, 5))
    src = dict(wavelength=1.9, te=0.6, tm=0.8, theta=17.0, phi=25.0)
    cl = build_crystal(st, eng, fields=True)
    cl.set_source(**src)
    cl.solve()
    E, H = cl.fields_volume(X, Y, z)
    kp = tuple(orc.kplanar(st["epsi"], src["wavelength"], src["theta"], src["phi"]))
    sol = orc.solve_structure(st, src["wavelength"], kp)
    Eo, Ho = orc.fields_volume(st, sol, X, Y, z, src["te"], src["tm"])
    assert np.abs(E - Eo).max() <= 1e-9 * np.abs(Eo).max()
    assert np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max()


@pytest.mark.parametrize("backend", BACKENDS)
def test_fields_grid_path_matches_point_path(backend):
    """Meshgrid coordinates take the separable transform (kh_fields_grid_batch); scattered points (here: the same grid
    sheared so that it is no longer a meshgrid, and a hexagonal lattice) take the dense phase matrix.  Same numbers."""
    eng = engine(backend)
    st, _, (X, Y, z) = cases.case_fields(5, slices=2, grid=(9, 7, 4))
    src = dict(wavelength=1.7, te=0.3, tm=0.9, theta=11.0, phi=40.0)
    cl = build_crystal(st, eng, fields=True)
    cl.set_source(**src)
    cl.solve()
    plan = cl._get_plan(True)
    assert plan.grid_separable
    E, H = cl.fields_volume(X, Y, z)                                       # grid path
    inc = np.hstack(cl.get_source_as_field_vectors()).reshape(1, 2, plan.n)
    F = eng.fields(plan, cl._solved, [cl.source.wavelength], [cl.kp], inc, X.ravel(), Y.ravel(), [float(v) for v in z],
                   cl.stack_positions).cpu().numpy()[0].reshape((len(z), 6) + X.shape)        # point path
    assert np.abs(E - F[:, :3]).max() <= 1e-12 * np.abs(F).max()
    assert np.abs(H - F[:, 3:]).max() <= 1e-12 * np.abs(F).max()
    Xs = X + 0.05 * Y                                                      # sheared: not a meshgrid -> point path inside the API
    E2, H2 = cl.fields_volume(Xs, Y, z)
    kp = tuple(orc.kplanar(st["epsi"], src["wavelength"], src["theta"], src["phi"]))
    sol = orc.solve_structure(st, src["wavelength"], kp)
    Eo, Ho = orc.fields_volume(st, sol, Xs, Y, z, src["te"], src["tm"])
    assert np.abs(E2 - Eo).max() <= 1e-9 * np.abs(Eo).max() and np.abs(H2 - Ho).max() <= 1e-9 * np.abs(Ho).max()


@pytest.mark.parametrize("backend", BACKENDS)
def test_fields_batched_groups_match_single_solves(backend):
    """The field pipeline batches its inverses over groups of stack positions / layers (up to 296 matrices per launch): a batch
    of 40 sources over the 14-position stack splits into two runs of 7 positions, depths that skip layers leave gaps between
    the runs.  Every solve of the batch must equal the same source solved alone (one group of all positions)."""
    eng = engine(backend)
    st, _, (X, Y, _) = cases.case_fields(3, slices=4, grid=(5, 4, 3))
    cl = build_crystal(st, eng, fields=True)
    B = 40
    wls = 1 / np.linspace(0.49, 0.6, B)
    cl.set_source(wavelength=float(wls[0]), te=1.0, tm=0.0)
    cl.solve()
    plan = cl._get_plan(True)
    assert plan.Ls == 14
    zp = np.asarray(cl.stack_positions)
    z = np.array([0.5 * (zp[1] + zp[2]), 0.3 * zp[2] + 0.7 * zp[3], 0.5 * (zp[5] + zp[6]), 0.5 * (zp[11] + zp[12])])     # positions 1, 2, 5, 11
    kp = np.zeros((B, 2), dtype=complex)
    pol = np.tile([[1.0, 0.0]], (B, 1)).astype(complex)
    inc = []
    for w in wls:
        cl.set_source(wavelength=float(w), te=1.0, tm=0.0)
        inc.append(np.hstack(cl.get_source_as_field_vectors()))
    inc = np.asarray(inc).reshape(B, 2, plan.n)
    solved = eng.solve_batch(plan, wls, kp, pol, want_flux=True, want_fields=True)
    F = eng.fields(plan, solved, wls, kp, inc, X.ravel(), Y.ravel(), z, zp).cpu().numpy()
    zall = np.linspace(0.01, zp[-2] - 0.01, 23)                                                  # every position active: runs of 7 + 7
    Fall = eng.fields(plan, solved, wls, kp, inc, X.ravel(), Y.ravel(), zall, zp).cpu().numpy()
    for b in (0, 17, 39):
        one = eng.solve_batch(plan, wls[b:b + 1], kp[b:b + 1], pol[b:b + 1], want_flux=True, want_fields=True)
        F1 = eng.fields(plan, one, wls[b:b + 1], kp[b:b + 1], inc[b:b + 1], X.ravel(), Y.ravel(), z, zp).cpu().numpy()[0]
        assert np.abs(F[b] - F1).max() <= 1e-11 * np.abs(F1).max()
        F1 = eng.fields(plan, one, wls[b:b + 1], kp[b:b + 1], inc[b:b + 1], X.ravel(), Y.ravel(), zall, zp).cpu().numpy()[0]
        assert np.abs(Fall[b] - F1).max() <= 1e-11 * np.abs(F1).max()


@pytest.mark.parametrize("backend", BACKENDS)
def test_bzi_beam_source_and_k_summed_fields(backend):
    """per-k source amplitudes (kh_beam_amplitudes) and the Brillouin-zone sum of the field maps, against the
    unmodified reference (examples/bzi/bzi_animation.py at test size: 1-D grating pw = (3, 1), 5 k-points)."""
    from khepri_b200 import Expansion
    from khepri_b200.beams import amplitudes_from_fields, bzi_fields, paraxial_gaussian_field, shifted_rotated_fields
    eng = engine(backend)
    g = gold("bzi_beam")
    st, c = cases.case_bzi_beam()
    b = c["beam"]
    X, Y = c["X"], c["Y"]
    src = shifted_rotated_fields(paraxial_gaussian_field, X, Y, np.zeros_like(X), b["wl"], b["x0"], b["y0"], b["z0"],
                                 b["theta"], b["phi"], b["pol"], beam_waist=b["beam_waist"], er=b["er"])
    src = np.swapaxes(np.swapaxes(np.asarray(src), 0, 2), 1, 3)
    assert np.abs(src - g["source"]).max() <= 1e-12 * np.abs(g["source"]).max()          # beam synthesis (host)
    e1 = Expansion(st["pw"])
    A = amplitudes_from_fields(g["source"], e1, c["wl"], c["kbz"], X, Y, c["bz"], engine=eng)       # [B, 4, N], batched over k
    assert A.shape == g["amplitudes"].shape
    assert np.abs(A - g["amplitudes"]).max() <= 1e-11 * np.abs(g["amplitudes"]).max()
    A0 = amplitudes_from_fields(g["source"], e1, c["wl"], tuple(c["kbz"][0]), X, Y, c["bz"], engine=eng)   # the reference's call
    assert A0.shape == (4, e1._g_vectors.shape[1]) and np.abs(A0 - A[0]).max() <= 1e-13 * np.abs(A).max()
    cl = build_crystal(st, eng, fields=True)
    xo, yo, zo = c["out"]
    E, H = bzi_fields(cl, c["wl"], c["kbz"], A.reshape(len(c["kbz"]), -1), xo, yo, zo)
    got = np.asarray((E, H))
    assert np.abs(got - g["fields"]).max() <= 1e-9 * np.abs(g["fields"]).max()


def _twisted(eng, tw, ta, fields=False):
    from khepri_b200 import Crystal, Expansion, Layer
    e1, e2 = Expansion(tw["pw"]), Expansion(tw["pw"])
    e1.rotate(ta / 2)
    e2.rotate(-ta / 2)
    cl = Crystal.from_expansion(e1 + e2, engine=eng)
    cl.add_layer("upper_layer", Layer.pixmap(e1, tw["pixmap"], tw["depths"][0]), extended=True)
    cl.add_layer("lower_layer", Layer.pixmap(e2, tw["pixmap"], tw["depths"][2]), extended=True)
    cl.add_layer("interlayer", Layer.uniform(e1, 1, tw["depths"][1]), extended=True)
    cl.set_device(["upper_layer", "interlayer", "lower_layer"], [fields] * 3)
    return cl


@pytest.mark.parametrize("backend", BACKENDS)
def test_twisted_bilayer_extended(backend):
    """notebooks/PRL_2021_BL.ipynb cells 2-8: (3,3)+(3,3) moire basis, n = 162."""
    eng = engine(backend)
    g = gold("twisted33")
    tw = cases.twisted_case()
    nt = len(tw["twists"]) if backend == "cuda" else 1
    for j in range(nt):
        cl = _twisted(eng, tw, tw["twists"][j])
        R, T = cl.solve_batch(1 / tw["freqs"], te=1, tm=0)
        assert np.abs(np.stack([R, T], 1) - g["RT"][:, j]).max() <= 1e-9
    cl.set_source(wavelength=1 / tw["freqs"][0], te=1, tm=0)
    cl.solve()
    assert np.abs(np.array(cl.poynting_flux_end()) - g["RT"][0, nt - 1]).max() <= 1e-9
    with pytest.raises(NotImplementedError):
        from khepri_b200 import Expansion, ExtendedLayer, Layer
        ExtendedLayer(cl.expansion, Layer.uniform(Expansion((3, 3)), 1, 0.1))     # base not part of the moire expansion


@pytest.mark.parametrize("backend", BACKENDS)
def test_twisted_bilayer_field_maps(backend):
    """Field maps of extended layers (extension.py:100-112: W, V, lambda of the shifted base solves scattered into the joint
    basis; crystal.py:250-253) against the unmodified reference: pixmap bases and the uniform spacer, normal and oblique source."""
    eng = engine(backend)
    g = gold("twisted33_fields")
    tw = cases.twisted_case()
    X, Y, z = cases.twisted_field_grid()
    todo = (("a", 1, 1, dict(te=1, tm=0)), ("b", 2, 0, dict(te=0.5, tm=1.0, theta=12.0, phi=20.0)))
    for tag, it, jf, src in (todo if backend == "cuda" else todo[:1]):
        cl = _twisted(eng, tw, tw["twists"][it], fields=True)
        cl.set_source(wavelength=1 / tw["freqs"][jf], **src)
        cl.solve()
        np.testing.assert_allclose(cl.poynting_flux_end(), g["RT_" + tag], rtol=1e-9)
        E, H = cl.fields_volume(X, Y, z)
        assert np.abs(E - g["E_" + tag]).max() <= 1e-9 * np.abs(g["E_" + tag]).max()
        assert np.abs(H - g["H_" + tag]).max() <= 1e-9 * np.abs(g["H_" + tag]).max()
        assert cl.layers["upper_layer"].W.shape == (162, 162) and cl.layers["interlayer"].L.shape == (162,)


# ----------------------------------------------------------------------------- sharding (gloo, world_size 2, CPU)
def test_shard_bounds_cover_everything():
    from khepri_b200.sharding import shard_bounds
    for total in (0, 1, 7, 151, 413696):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(total, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def _rank_main(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from khepri_b200.sharding import allreduce_sum, sweep_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = engine("emu")
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng)
    wl = np.array([s["wavelength"] for s in srcs[:7]])
    R, T = sweep_sharded(cl, wl, te=1.0, tm=0.0)
    part = torch.full((3,), complex(rank + 1, -rank), dtype=torch.complex128)
    tot = allreduce_sum(part)
    # Brillouin-zone-integrated field maps: 5 k-points over 2 ranks (3 + 2), partial sums all-reduced
    from khepri_b200.beams import bzi_fields
    stb, c = cases.case_bzi_beam()
    clb = build_crystal(stb, eng, fields=True)
    xo, yo, zo = c["out"]
    E, H = bzi_fields(clb, c["wl"], c["kbz"], gold("bzi_beam")["amplitudes"].reshape(len(c["kbz"]), -1), xo, yo, zo)
    if rank == 0:
        np.savez(out, R=R, T=T, tot=tot.numpy(), EH=np.asarray((E, H)))
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_sharded_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    res = np.load(out)
    g = gold("suh03")
    assert np.abs(np.stack([res["R"], res["T"]], 1) - g["RT"][:7]).max() <= 1e-9
    assert np.allclose(res["tot"], complex(3, -1))
    gb = gold("bzi_beam")["fields"]
    assert np.abs(res["EH"] - gb).max() <= 1e-9 * np.abs(gb).max()


@pytest.mark.gpu
def test_sweep_sharded_single_gpu_passthrough():
    from khepri_b200.sharding import sweep_sharded
    eng = engine("cuda")
    st, srcs = cases.case_suh03()
    cl = build_crystal(st, eng)
    wl = np.array([s["wavelength"] for s in srcs])
    R, T = sweep_sharded(cl, wl, te=1.0, tm=0.0)
    assert np.abs(np.stack([R, T], 1) - gold("suh03")["RT"]).max() <= 1e-9
