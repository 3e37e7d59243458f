"""Truncated SVD of a device-resident form-factor block by a randomised range
finder (Halko, Martinsson, Tropp 2011) -- the device feed for the SVD leaves of
the reference's ``CompressedFormFactorMatrix``.

The reference computes the same factors with ARPACK ``svds`` on the CPU, one
matrix-vector product at a time (src/flux/linalg.py:8-50, called from
src/flux/compressed_form_factors.py:388-405).  Here the products with the block
run on the slab in HBM (``csr_matmat_kernel`` / ``csr_rmatmat_kernel``); the
small dense factorizations (QR / SVD of k-column matrices) are torch.linalg
library calls on the device.
"""
import numpy as np


def sparse_svd(block, k, oversample=10, n_iter=4, seed=0):
    """Leading ``k`` singular triplets ``(U, S, Vt)`` of a
    :class:`~fluxpy_b200.device_csr.DeviceCsrSlab` -- same return convention as
    ``flux.linalg.sparse_svd`` (src/flux/linalg.py:8-18): NumPy arrays, singular
    values in DESCENDING order."""
    import torch
    m, n = block.shape
    dev = torch.device('cuda', block.device)
    p = int(min(min(m, n), k + oversample))
    if p == 0:
        return np.zeros((m, 0)), np.zeros(0), np.zeros((0, n))
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Q = torch.randn(n, p, dtype=torch.float64, device=dev, generator=g)
    Q, _ = torch.linalg.qr(block.matmat(Q))                  # range of A
    for _ in range(n_iter):                                  # power iterations, re-orthonormalised
        Z, _ = torch.linalg.qr(block.rmatmat(Q))
        Q, _ = torch.linalg.qr(block.matmat(Z))
    B = block.rmatmat(Q).T                                   # p x n  (= Q^T A)
    Ub, S, Vt = torch.linalg.svd(B, full_matrices=False)
    U = Q@Ub
    return U[:, :k].cpu().numpy(), S[:k].cpu().numpy(), Vt[:k].cpu().numpy()


def estimate_rank(block, tol, max_nbytes=None, k0=40):
    """Smallest ``k`` whose next singular value is below ``tol`` times the
    largest, by doubling ``k`` as ``flux.linalg.estimate_rank`` does
    (src/flux/linalg.py:20-50); returns ``(U, S, Vt, tol_reached)`` or ``None``
    when the factors would need more than ``max_nbytes``."""
    m, n = block.shape
    k = min(k0, min(m, n))
    while True:
        U, S, Vt = sparse_svd(block, k)
        if S.size == 0:
            return U, S, Vt, 0.0
        small = np.where(S < tol*S[0])[0]
        if small.size or k >= min(m, n):
            r = int(small[0]) if small.size else S.size
            r = max(r, 1)
            if max_nbytes is not None and (U[:, :r].nbytes + S[:r].nbytes + Vt[:r].nbytes) > max_nbytes:
                return None
            return U[:, :r], S[:r], Vt[:r], float(S[r]/S[0]) if r < S.size else 0.0
        if max_nbytes is not None and 8*(m + n + 1)*2*k > max_nbytes:
            return None
        k = min(2*k, min(m, n))
