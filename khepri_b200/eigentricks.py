"""Band-structure post-processing of a unit-cell S-matrix (khepri/eigentricks.py:5-60), batched on the GPU.

The reference solves the generalized problem  Sl v = w Sr v  with scipy's QZ (eigentricks.py:29-40), where
(p = forward half, m = backward half of the 2n x 2n block matrix)

    Sl = [[S_pp, 0], [S_mp, -I]],      Sr = [[I, -S_pm], [0, -S_mm]]        (eigentricks.py:5-21).

Sr is block triangular, so Sr^-1 Sl has the closed form (the transfer matrix of the cell)

    M = [[S_pp - S_pm S_mm^-1 S_mp,  S_pm S_mm^-1], [-S_mm^-1 S_mp,  S_mm^-1]]

with the same eigenpairs (w, v).  It is built for a whole batch of S-matrices with one batched inverse and three batched
DMMA GEMMs and handed to the batched non-Hermitian eigensolver (the same kernels as the RCWA solve).  The Bloch factors
on the unit circle (``on_shell``) -- the ones band diagrams keep -- are well conditioned in this form; eigenvalues that
are infinite in the pencil (singular S_mm) come out huge instead of ``inf``.
"""
import numpy as np
import torch

from .engine import Engine


def _blocks(S):
    S = np.asarray(S) if not torch.is_tensor(S) else S
    if S.ndim >= 4 and S.shape[-4:-2] == (2, 2):                       # Crystal.Stot layout (..., 2, 2, n, n)
        lead = S.shape[:-4]
        return lead, S[..., 0, 0, :, :], S[..., 0, 1, :, :], S[..., 1, 0, :, :], S[..., 1, 1, :, :]
    h = S.shape[-1] // 2                                               # flat block matrix (..., 2n, 2n) as in the reference
    return S.shape[:-2], S[..., :h, :h], S[..., :h, h:], S[..., h:, :h], S[..., h:, h:]


def scattering_splitlr(S):
    """Pencil (Sl, Sr) of a flat 2n x 2n S-matrix on the host (eigentricks.py:5-21); kept for callers that run their own eig."""
    S = np.asarray(S)
    h = S.shape[0] // 2
    Sl, Sr = np.zeros_like(S), np.zeros_like(S)
    Sl[:h, :h] = S[:h, :h]; Sl[h:, :h] = S[h:, :h]; Sl[h:, h:] = -np.eye(h, dtype=S.dtype)
    Sr[:h, :h] = np.eye(h, dtype=S.dtype); Sr[:h, h:] = -S[:h, h:]; Sr[h:, h:] = -S[h:, h:]
    return Sl, Sr


def _transfer(eng, Spp, Spm, Smp, Smm):
    """[B, 2n, 2n] transfer matrices on the device."""
    f = lambda x: eng.to_dev(np.ascontiguousarray(x) if not torch.is_tensor(x) else x.contiguous(), torch.complex128)
    Spp, Spm, Smp, Smm = f(Spp), f(Spm), f(Smp), f(Smm)
    B, n, _ = Smm.shape
    Xi, info = eng.zinv(Smm, return_info=True)                         # S_mm^-1
    T12 = eng.zgemm(Spm, Xi)                                           # S_pm S_mm^-1
    T21 = eng.zgemm(Xi, Smp, alpha=-1.0)                               # -S_mm^-1 S_mp
    T11 = Spp + eng.zgemm(Spm, T21)                                    # S_pp - S_pm S_mm^-1 S_mp
    M = torch.empty((B, 2 * n, 2 * n), dtype=torch.complex128, device=Xi.device)
    M[:, :n, :n] = T11; M[:, :n, n:] = T12; M[:, n:, :n] = T21; M[:, n:, n:] = Xi
    return M, info


def scattering_eigenvalues(S, dos=False, engine=None):
    """(w, v[, det]) of the pencil of eigentricks.py:29-40 for one S-matrix or a batch (leading axes), computed on the GPU.

    ``S`` is ``Crystal.Stot`` (..., 2, 2, n, n) or the reference's flat (..., 2n, 2n) block matrix.  Returns NumPy arrays
    ``w`` (..., 2n) and ``v`` (..., 2n, 2n) (eigenvectors in columns, arbitrary order and scale as with LAPACK); ``None``
    if S holds NaNs (eigentricks.py:33-34)."""
    eng = engine or Engine.default()
    lead, Spp, Spm, Smp, Smm = _blocks(S)
    n = Smm.shape[-1]
    rs = lambda x: x.reshape((-1, n, n))
    Spp, Spm, Smp, Smm = rs(Spp), rs(Spm), rs(Smp), rs(Smm)
    nan = any(bool(torch.isnan(torch.view_as_real(x)).any()) if torch.is_tensor(x) else bool(np.isnan(x).any()) for x in (Spp, Spm, Smp, Smm))
    if nan:
        return None
    M, _ = _transfer(eng, Spp, Spm, Smp, Smm)
    w, v, _ = eng.zgeev(M)
    w = w.cpu().numpy().reshape(lead + (2 * n,))
    v = v.cpu().numpy().reshape(lead + (2 * n, 2 * n))
    if not dos:
        return w, v
    return w, v, scattering_det(S, engine=eng)


def scattering_det(S, engine=None):
    """det(S_pp - S_pm S_mm^-1 S_mp) det(S_mm)  (eigentricks.py:23-27), as products of eigenvalues from the batched eigensolver."""
    eng = engine or Engine.default()
    lead, Spp, Spm, Smp, Smm = _blocks(S)
    n = Smm.shape[-1]
    rs = lambda x: x.reshape((-1, n, n))
    M, _ = _transfer(eng, rs(Spp), rs(Spm), rs(Smp), rs(Smm))
    w1, _, _ = eng.zgeev(M[:, :n, :n].contiguous())
    w2, _, _ = eng.zgeev(eng.to_dev(np.ascontiguousarray(rs(Smm)) if not torch.is_tensor(Smm) else rs(Smm).contiguous(), torch.complex128))
    d = (torch.prod(w1, dim=1) * torch.prod(w2, dim=1)).cpu().numpy().reshape(lead)
    return d if lead else complex(d)


def on_shell(eigenvalues, tol=1e-10):
    """|w| = 1 within tol (eigentricks.py:42-43)."""
    return np.isclose(np.abs(eigenvalues), 1.0, rtol=0.0, atol=tol)


def band_structure(S, engine=None):
    """eigentricks.py:47-56 (the reference keeps the REAL parts in (0, 1) that are on shell)."""
    res = scattering_eigenvalues(S, engine=engine)
    if res is None:
        return None
    w = res[0].real
    w = w[w > 0]
    w = w[w < 1]
    return w[on_shell(w)]
