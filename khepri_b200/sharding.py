"""Multi-GPU sharding of a sweep.

Every (frequency, k-point) solve is independent, so the flattened batch is partitioned contiguously
over the ranks (one process per GPU) with NO data-path collective; torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) is used only to all-gather the flux spectra and to
sum Brillouin-zone-integrated field maps.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous, balanced partition: the first (total % world) ranks take one extra item."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_spectra(local, total, group=None):
    """All-gather per-rank [B_local, ...] float64 spectra into the full [total, ...] array (every rank gets it)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return full if isinstance(local, torch.Tensor) else full.cpu().numpy()


def allreduce_sum(x, group=None):
    """Sum of Brillouin-zone partial field maps over the ranks (complex tensors go as float pairs)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    view = torch.view_as_real(t) if t.is_complex() else t
    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
    return t if isinstance(x, torch.Tensor) else t.cpu().numpy()


def sweep_sharded(crystal, wavelengths, kps=None, te=1.0, tm=1.0, theta=0.0, phi=0.0, group=None):
    """Crystal.solve_batch over this rank's contiguous shard, then all-gather: every rank returns the
    full (R[B], T[B])."""
    wl = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64)).reshape(-1)
    B = wl.size
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(B, world, rank)

    def cut(a):
        a = np.asarray(a)
        return a if a.ndim == 0 else a[lo:hi]

    R, T = crystal.solve_batch(wl[lo:hi], kps=None if kps is None else np.asarray(kps)[lo:hi],
                               te=cut(te), tm=cut(tm), theta=cut(theta), phi=cut(phi))
    full = gather_spectra(np.stack([R, T], axis=1), B, group)
    return full[:, 0], full[:, 1]
