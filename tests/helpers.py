"""Shared checkers for the parity tests (CPU and GPU tiers)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def unit(dtype):
    return float(np.finfo(dtype).eps)/2


def numerator_band(P, N, I, J, dtype):
    """Bound on the round-off of the reference's numerator
    ``max(0, N[i]@(P[J]-P[i]).T) * max(0, P[i]@N[J].T - NJ_PJ)``
    (form_factors.py:46-47) evaluated in ``dtype``: pairs whose exact numerator
    lies within this band of ``eps`` may legitimately fall on either side of
    the cull.  Dense (len(I), len(J)) float64."""
    u = unit(dtype)
    P = np.asarray(P, np.float64)
    N = np.asarray(N, np.float64)
    d = P[J][None] - P[I][:, None]
    a = np.abs(np.einsum('ik,ijk->ij', N[I], d))
    b = np.abs(np.einsum('jk,ijk->ij', N[J], d))
    dn = np.sqrt((d*d).sum(-1))
    pn = np.sqrt((P*P).sum(1))
    err_a = 8*u*(dn + pn[I][:, None])                      # N[i]@(P[J]-P[i])
    err_b = 8*u*(pn[I][:, None] + pn[J][None, :])          # P[i]@N[J] - sum(N[J]*P[J])
    return a*err_b + b*err_a + err_a*err_b + 4*u*a*b


def check_against_reference_csr(FF, ref_indptr, ref_indices, ref_data, P, N, A, I, J, eps,
                                dtype, vis=None, rtol=None):
    """FF (ours, direct evaluation) against a CSR produced by the reference's
    own Python in ``dtype``.

    * pattern: identical except for pairs inside ``numerator_band`` of eps;
    * values on the common pattern: within the reference's own round-off of the
      float64 ground truth, and ours within ``rtol`` of that ground truth.
    """
    from oracle.oracle import form_factor_dense_f64
    m, n = len(I), len(J)
    ours = np.zeros((m, n), bool)
    ref = np.zeros((m, n), bool)
    FF = FF.tocsr()
    FF.sort_indices()
    rows = np.repeat(np.arange(m), np.diff(FF.indptr))
    ours[rows, FF.indices] = True
    rrows = np.repeat(np.arange(m), np.diff(ref_indptr))
    ref[rrows, ref_indices] = True
    num, val = form_factor_dense_f64(P, N, A, I, J)
    diff = ours != ref
    if diff.any():
        band = numerator_band(P, N, I, J, dtype)
        assert (np.abs(num - eps)[diff] <= band[diff] + unit(dtype)*eps).all(), \
            'pattern differs outside the cull round-off band'
    common = ours & ref
    ours_val = np.zeros((m, n))
    ours_val[rows, FF.indices] = FF.data
    ref_val = np.zeros((m, n))
    ref_val[rrows, ref_indices] = ref_data
    if rtol is None:
        rtol = 1e-5 if dtype == np.float32 else 1e-12
    t = val[common]
    assert (np.abs(ours_val[common] - t) <= rtol*np.abs(t)).all(), 'value outside rtol of f64 truth'
    # the reference's own round-off on the same entries (for the record / sanity)
    band = numerator_band(P, N, I, J, dtype)[common]
    ok = np.abs(ref_val[common] - t) <= (band/np.maximum(num[common], 1e-300) + 16*unit(dtype))*np.abs(t)
    assert ok.all(), 'reference value outside its own round-off bound: checker is wrong'
    return int(diff.sum())
