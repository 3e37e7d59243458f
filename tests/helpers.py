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


def heightfield_clearance(V, n, pi, PJ, ray_offset=1e-3):
    """Independent geometric ground truth for centroid-to-centroid visibility on a regular-grid height field
    triangulated as ``fluxpy_b200.meshes.grid_faces`` (cells split along the a-d diagonal) -- no ray/triangle
    test, no BVH, no oracle code.  The surface is piecewise linear, so along the segment p_i -> p_j the
    clearance g(s) = z_segment(s) - z_surface(x(s), y(s)) is piecewise linear with breakpoints where the
    segment's xy-projection crosses a grid line x = x_k, y = y_k or a cell diagonal; the segment crosses the
    surface (some triangle is hit before the target, from either side: shape.py:349-398 has no back-face
    culling) iff g changes sign over the breakpoints.  Returns (gmin, gmax) over the breakpoints of every
    target (+inf / -inf when there is none): visible iff gmin > 0 or gmax < 0; min(|gmin|, |gmax|) small =
    a grazing pair that floating-point ray tracers may legitimately call either way."""
    V = np.asarray(V, np.float64)
    Z = V[:, 2].reshape(n, n)                   # Z[iy, ix]
    x0, y0 = V[0, 0], V[0, 1]
    h = (V[n*n - 1, 0] - x0)/(n - 1)
    pi = np.asarray(pi, np.float64)
    PJ = np.asarray(PJ, np.float64)
    uA, vA, zA = (pi[0] - x0)/h, (pi[1] - y0)/h, pi[2]
    uB, vB, zB = (PJ[:, 0] - x0)/h, (PJ[:, 1] - y0)/h, PJ[:, 2]
    L = np.sqrt(((PJ - pi)**2).sum(1))
    gmin = np.full(len(PJ), np.inf)
    gmax = np.full(len(PJ), -np.inf)

    def lerp_cell(iy0, ix0, iy1, ix1, f):
        return Z[iy0, ix0]*(1 - f) + Z[iy1, ix1]*f

    def account(s, height, ok):
        nonlocal gmin, gmax
        ok = ok & (s > 0) & (s < 1) & (s*L[:, None] > ray_offset)
        g = zA + (zB - zA)[:, None]*s - height
        gmin = np.minimum(gmin, np.where(ok, g, np.inf).min(1))
        gmax = np.maximum(gmax, np.where(ok, g, -np.inf).max(1))

    with np.errstate(divide='ignore', invalid='ignore'):
        k = np.arange(n, dtype=np.float64)[None, :]
        # grid lines u = k: the surface along the line is linear between the nodes (k, floor v), (k, floor v + 1)
        s = (k - uA)/(uB - uA)[:, None]
        ok = np.isfinite(s)
        s = np.where(ok, s, 0.5)
        v = vA + (vB - vA)[:, None]*s
        c = np.clip(np.floor(v), 0, n - 2).astype(int)
        kk = np.broadcast_to(k.astype(int), c.shape)
        account(s, lerp_cell(c, kk, c + 1, kk, v - c), ok)
        # grid lines v = k
        s = (k - vA)/(vB - vA)[:, None]
        ok = np.isfinite(s)
        s = np.where(ok, s, 0.5)
        u = uA + (uB - uA)[:, None]*s
        c = np.clip(np.floor(u), 0, n - 2).astype(int)
        account(s, lerp_cell(kk, c, kk, c + 1, u - c), ok)
        # cell diagonals u - v = k: between the nodes (floor u, floor v) and (floor u + 1, floor v + 1)
        k = np.arange(-(n - 1), n, dtype=np.float64)[None, :]
        wA, wB = uA - vA, uB - vB
        s = (k - wA)/(wB - wA)[:, None]
        ok = np.isfinite(s)
        s = np.where(ok, s, 0.5)
        u = uA + (uB - uA)[:, None]*s
        v = u - k                                # exactly on the diagonal
        cu = np.clip(np.floor(u), 0, n - 2).astype(int)
        cv = np.clip(cu - k.astype(int), 0, n - 2)
        ok = ok & (cu - k.astype(int) >= 0) & (cu - k.astype(int) <= n - 2)
        account(s, lerp_cell(cv, cu, cv + 1, cu + 1, u - cu), ok)
    return gmin, gmax


def heightfield_sun_clearance(V, n, P, N, D, eps=None):
    """Same geometry for the sun-occlusion query (shape.py:400-421): the ray from P[i] + 1e-3*N[i] along D,
    followed to where it leaves the mesh's xy bounding box; blocked iff its clearance over the surface is
    negative at some grid-line / diagonal crossing.  Returns gmin per face (+inf: no crossing inside)."""
    V = np.asarray(V, np.float64)
    P = np.asarray(P, np.float64)
    N = np.asarray(N, np.float64)
    D = np.asarray(D, np.float64)
    if eps is None:
        eps = 1e3*np.finfo(np.float32).resolution
    lo, hi = V[:, :2].min(0), V[:, :2].max(0)
    out = np.empty(len(P))
    for i in range(len(P)):
        org = P[i] + eps*N[i]
        with np.errstate(divide='ignore'):
            t_exit = np.where(D[:2] > 0, (hi - org[:2])/D[:2], np.where(D[:2] < 0, (lo - org[:2])/D[:2], np.inf)).min()
        end = org + (t_exit + 1e-6)*D            # a hair past the boundary, so that the crossing of the mesh's last grid line counts
        gmin, _ = heightfield_clearance(V, n, org, end[None, :], ray_offset=0.0)
        out[i] = gmin[0]
    return out
