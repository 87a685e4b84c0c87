"""Brillouin-zone-integration sources (khepri/beams.py).  Beam synthesis is input generation on the host; the per-k
Fourier amplitudes (amplitudes_from_fields, beams.py:164-191) and the k-sum of the field maps -- the steps either side
of the batched solve in examples/bzi/*.py -- run on the GPU: one DMMA GEMM per k-point for the amplitudes, the
retained-eigenspace solve + field reconstruction batched over the k-points, a device sum over k and (multi-GPU) one
all-reduce of the summed maps."""
from math import prod

import numpy as np
import torch

from .sharding import allreduce_sum, shard_bounds


def gen_bzi_grid(shape, a=1, reciproc=None):
    """Midpoints of a shape[0] x shape[1] partition of the first Brillouin zone (beams.py:193-207) -> (2, n0, n1)."""
    b1, b2 = ([2 * np.pi / a, 0], [0, 2 * np.pi / a]) if reciproc is None else reciproc
    si, sj = 1 / shape[0], 1 / shape[1]
    i, j = np.meshgrid(np.arange(-0.5 + si / 2, 0.5, si), np.arange(-0.5 + sj / 2, 0.5, sj), indexing="ij")
    return np.stack([b1[0] * i + b2[0] * j, b1[1] * i + b2[1] * j])


def rotation_matrix(polar_angle, azimuthal_angle, polarization_angle):
    """beams.py:6-47: R_p(axis = propagation direction) R_z(azimuth) R_y(polar)."""
    cp, sp = np.cos(polar_angle), np.sin(polar_angle)
    ca, sa = np.cos(azimuthal_angle), np.sin(azimuthal_angle)
    co, so = np.cos(polarization_angle), np.sin(polarization_angle)
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    rz = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]])
    u = np.array([ca * sp, sa * sp, cp])
    ux = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    rp = co * np.eye(3) + (1 - co) * np.outer(u, u) + so * ux          # Rodrigues
    return rp @ rz @ ry


def paraxial_gaussian_field(x, y, z, wl, beam_waist=1, er=1):
    """beams.py:136-160: paraxial Gaussian beam polarised along x, H = E / sqrt(er) along y."""
    k = 2 * np.pi / wl
    z_r = np.pi * beam_waist ** 2 * np.sqrt(er) / wl
    w_z = beam_waist * np.sqrt(1 + (z / z_r) ** 2)
    r2 = x ** 2 + y ** 2
    ex = beam_waist / w_z * np.exp(-r2 / w_z ** 2) * np.exp(1j * (k * z + k * r2 / 2 * z / (z ** 2 + z_r ** 2) - np.arctan(z / z_r)))
    zero = np.zeros_like(ex)
    return (ex, zero, zero), (zero, ex / np.sqrt(er), zero)


_paraxial_gaussian_field_fn = paraxial_gaussian_field


def shifted_rotated_fields(field_fn, x, y, z, wavelength, beam_origin_x, beam_origin_y, beam_origin_z,
                           polar_angle, azimuthal_angle, polarization_angle, **kwargs):
    """beams.py:50-101: evaluate field_fn in the rotated, shifted beam frame and rotate the vectors back -> (2, 3, *x.shape)."""
    mat = rotation_matrix(polar_angle, azimuthal_angle, polarization_angle)
    inv = np.linalg.inv(mat)
    pts = np.stack([x, y, z], axis=-1)
    rot = pts @ inv.T
    o = inv @ np.array([beam_origin_x, beam_origin_y, beam_origin_z], dtype=float)
    e, h = field_fn(rot[..., 0] - o[0], rot[..., 1] - o[1], rot[..., 2] - o[2], wavelength, **kwargs)
    ef = np.stack(e, axis=-1) @ mat.T
    hf = np.stack(h, axis=-1) @ mat.T
    return np.asarray([tuple(ef[..., i] for i in range(3)), tuple(hf[..., i] for i in range(3))])


def amplitudes_from_fields(fields, e, wl, kp, x, y, bzs, a=1, engine=None):
    """Fourier amplitudes (Ex, Ey, Hx, Hy)_g of a real-space source for one k-point -- or a whole array kp[B, 2] of them
    (beams.py:164-191).  fields: (ny, nx, 2, 3) samples at the supercell points (x, y) (2-D meshgrids covering bzs[0] x
    bzs[1] unit cells); returns (4, N) like the reference for a single k-point, (B, 4, N) for a batch."""
    from .engine import Engine
    eng = engine if engine is not None else Engine.default()
    fields = np.asarray(fields)
    ny, nx = fields.shape[:2]
    assert nx % bzs[0] == 0 and ny % bzs[1] == 0 and nx // bzs[0] == ny // bzs[1], "square tiles of NS x NS samples per unit cell"
    NS = ny // bzs[1]
    scale = 1.0 / (prod(bzs) * NS)                                  # the reference's normalisation: / n_tiles / NS
    kps = np.asarray(kp)
    single = kps.ndim == 1
    kps = np.atleast_2d(kps).astype(np.complex128)
    F4 = np.ascontiguousarray(fields[..., :2].reshape(ny * nx, 4), dtype=np.complex128)       # (Ex, Ey, Hx, Hy) per sample
    amp = eng.beam_amplitudes(kps, e._g_vectors, np.asarray(x, float).reshape(-1), np.asarray(y, float).reshape(-1), F4, scale)
    out = amp.transpose(1, 2).cpu().numpy()                          # [B, 4, N]
    return out[0] if single else out


def bzi_fields(crystal, wavelength, kps, incident, x, y, z, group=None):
    """Brillouin-zone-integrated field maps (examples/bzi/bzi_animation.py:59-80): for every k-point solve the stack with
    retained eigenspaces, reconstruct E, H on (x, y, z) from that k-point's incident amplitudes and SUM over k.  The
    k-points are sharded over the ranks of `group` (one process per GPU); the partial sums are all-reduced.
    incident: [B, 4N] or [B, 4, N] amplitudes from amplitudes_from_fields.  Returns (E, H), each (nz, 3, ny, nx)."""
    import torch.distributed as dist
    kps = np.asarray(kps).reshape(-1, 2)
    B = kps.shape[0]
    inc = np.asarray(incident, dtype=np.complex128).reshape(B, -1)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(B, world, rank)
    x, y = np.asarray(x, float), np.asarray(y, float)
    zs = [float(v) for v in np.atleast_1d(z)]
    total = crystal.fields_batch_sum([wavelength] * (hi - lo), kps[lo:hi], inc[lo:hi], x, y, zs)
    total = allreduce_sum(total, group)
    F = total.cpu().numpy().reshape((len(zs), 6) + x.shape)
    return F[:, :3], F[:, 3:]
