"""CPU oracle for the form-factor assembly path -- TEST INFRASTRUCTURE ONLY.

Python face of ``oracle/ff_oracle.c`` (see that file's header for the parity
status and the arithmetic contract).  Nothing under ``fluxpy_b200/`` may
import this module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Restated reference functions (citations relative to /root/reference):

* ``get_form_factor_matrix``            src/flux/form_factors.py:11-72
* ``EmbreeTrimeshShapeModel`` hooks     src/flux/shape.py:295-421
* face geometry helpers                 src/flux/shape.py:16-45
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libff_oracle.so')

DEFAULT_EPS = 1e-5  # src/flux/config.py:2


def build(force=False):
    """Compile the C oracle with gcc (idempotent)."""
    src = os.path.join(_HERE, 'ff_oracle.c')
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    subprocess.check_call(['make', '-s', '-C', _HERE, 'all'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = ctypes.CDLL(_SO)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.oracle_ray_eps.restype = ctypes.c_float
    L.oracle_scene_create.restype = vp
    L.oracle_scene_create.argtypes = [vp, sz, vp, sz]
    L.oracle_scene_destroy.argtypes = [vp]
    L.oracle_intersect1M.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, i32, i32]
    L.oracle_occluded1M.argtypes = [vp, sz, vp, vp, vp, vp, i32, i32]
    for sfx in ('f32', 'f64'):
        getattr(L, 'oracle_setup_ray_' + sfx).argtypes = [vp, vp, vp, vp]
        getattr(L, 'oracle_setup_ray_' + sfx).restype = i32
        getattr(L, 'oracle_visibility_' + sfx).argtypes = [vp, vp, vp, sz, vp, sz, vp, i32, i32]
        getattr(L, 'oracle_is_occluded_' + sfx).argtypes = [vp, vp, vp, vp, sz, vp, sz, i32, vp, i32, i32]
        f = getattr(L, 'oracle_ff_assemble_' + sfx)
        f.restype = vp
        f.argtypes = [vp, vp, vp, vp, vp, sz, vp, sz, ctypes.c_double, i32, i32]
        getattr(L, 'oracle_face_geometry_' + sfx).argtypes = [vp, vp, sz, vp, vp, vp]
    L.oracle_team_size.argtypes = [i32]
    L.oracle_set_study_variant.argtypes = [i32, i32]
    L.oracle_team_size.restype = i32
    L.oracle_ff_nnz.restype = ctypes.c_int64
    L.oracle_ff_nnz.argtypes = [vp]
    L.oracle_ff_tested.restype = ctypes.c_int64
    L.oracle_ff_tested.argtypes = [vp]
    L.oracle_ff_copy.argtypes = [vp, vp, vp, vp]
    L.oracle_ff_free.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dtype):
    if dtype == np.float32:
        return 'f32'
    if dtype == np.float64:
        return 'f64'
    raise RuntimeError(f'unsupported dtype {dtype}')  # form_factors.py:37


def face_geometry(V, F):
    """P, N, A as ``flux.shape`` computes them (shape.py:16-45), in ``V.dtype``."""
    V = np.ascontiguousarray(V)
    F64 = np.ascontiguousarray(F, dtype=np.int64)
    nf = F64.shape[0]
    P = np.empty((nf, 3), V.dtype)
    N = np.empty((nf, 3), V.dtype)
    A = np.empty((nf,), V.dtype)
    getattr(lib(), 'oracle_face_geometry_' + _sfx(V.dtype))(
        _ptr(V), _ptr(F64), nf, _ptr(P), _ptr(N), _ptr(A))
    return P, N, A


def team_size(nthreads=0):
    """OpenMP threads a region asked for ``nthreads`` (0 = default) really gets."""
    return int(lib().oracle_team_size(int(nthreads)))


def set_study_variant(tmode=0, tie=0):
    """Study switches of ff_oracle.c (hit distance as Embree's rcp + Newton step, equal-t winner by smallest
    index).  (0, 0) is the contract; anything else is for the sensitivity tests only."""
    lib().oracle_set_study_variant(int(tmode), int(tie))


class OracleScene:
    """float32 vertex buffer + uint32 index buffer, as shape.py:319-333 hands
    them to Embree."""

    def __init__(self, V, F):
        self._V32 = np.ascontiguousarray(V, dtype=np.float32)
        self._F32 = np.ascontiguousarray(F, dtype=np.uint32)
        self.nf = self._F32.shape[0]
        self._h = lib().oracle_scene_create(
            _ptr(self._V32), self._V32.shape[0], _ptr(self._F32), self.nf)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.oracle_scene_destroy(h)

    def intersect1M(self, org, dir, tnear, tfar, prim_id, geom_id, use_bvh=True, nthreads=0):
        n = org.shape[0]
        for a, dt in ((org, np.float32), (dir, np.float32), (tnear, np.float32),
                      (tfar, np.float32), (prim_id, np.uint32), (geom_id, np.uint32)):
            assert a.dtype == dt and a.flags.c_contiguous
        lib().oracle_intersect1M(self._h, n, _ptr(org), _ptr(dir), _ptr(tnear), _ptr(tfar),
                                 _ptr(prim_id), _ptr(geom_id), int(use_bvh), nthreads)

    def occluded1M(self, org, dir, tnear, tfar, use_bvh=True, nthreads=0):
        n = org.shape[0]
        for a in (org, dir, tnear, tfar):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        lib().oracle_occluded1M(self._h, n, _ptr(org), _ptr(dir), _ptr(tnear), _ptr(tfar),
                                int(use_bvh), nthreads)


class OracleShapeModel:
    """The slice of ``EmbreeTrimeshShapeModel`` that the hot path touches.

    Attributes follow ``TrimeshShapeModel`` (shape.py:55-127); ``N`` may be
    mutated in place by the caller, exactly as the reference tests do
    (tests/test_form_factors.py:33-34).
    """

    def __init__(self, V, F, N=None, P=None, A=None, use_bvh=True, nthreads=0):
        self.dtype = V.dtype
        _sfx(self.dtype)
        self.V, self.F = V, F
        P0, N0, A0 = face_geometry(V, F)
        self.P = P0  # the reference recomputes P regardless (shape.py:104)
        self.N = N0 if N is None else N
        self.A = A0 if A is None else A
        self.use_bvh, self.nthreads = use_bvh, nthreads
        self.scene = OracleScene(V, F)

    @property
    def num_faces(self):
        return self.P.shape[0]

    def _idx(self, I):
        return np.ascontiguousarray(np.asarray(I).astype(np.uint64))

    def get_visibility(self, I, J, oriented=False):
        I, J = self._idx(I), self._idx(J)
        P = np.ascontiguousarray(self.P)
        vis = np.empty((len(I), len(J)), np.uint8)
        getattr(lib(), 'oracle_visibility_' + _sfx(self.dtype))(
            self.scene._h, _ptr(P), _ptr(I), len(I), _ptr(J), len(J), _ptr(vis),
            int(self.use_bvh), self.nthreads)
        vis = vis.astype(bool)
        if oriented:  # shape.py:157-161
            Ii, Jj = np.where(vis)
            gi, gj = I[Ii].astype(np.int64), J[Jj].astype(np.int64)
            bad = ((self.P[gj] - self.P[gi]) * self.N[gi]).sum(1) <= 0
            vis[Ii[bad], Jj[bad]] = False
        return vis

    def get_visibility_1_to_N(self, i, J, oriented=False):
        return self.get_visibility([i], J, oriented).ravel()

    def get_visibility_matrix(self, oriented=False):
        I = np.arange(self.num_faces, dtype=np.uintp)
        return self.get_visibility(I, I, oriented)

    def is_occluded(self, I, D):
        I = self._idx(I)
        D = np.ascontiguousarray(D, dtype=self.dtype)
        if D.ndim not in (1, 2):
            raise ValueError('D.ndim should be 1 or 2')
        P = np.ascontiguousarray(self.P)
        N = np.ascontiguousarray(self.N)
        m = len(I)
        # Embree's ray.dir[:] = D broadcasts a (3,) vector to every ray and needs
        # D.shape == (m, 3) otherwise (shape.py:411)
        if D.ndim == 1:
            nd, per_face = 1, 0
        else:
            if D.shape[0] != m:
                raise ValueError('need D.shape[0] == len(I)')
            nd, per_face = m, 1
        occ = np.empty((m,), np.uint8)
        getattr(lib(), 'oracle_is_occluded_' + _sfx(self.dtype))(
            self.scene._h, _ptr(P), _ptr(N), _ptr(I), m, _ptr(D), nd, per_face, _ptr(occ),
            int(self.use_bvh), self.nthreads)
        return occ.astype(bool)


def get_form_factor_matrix(shape_model, I=None, J=None, eps=None, return_stats=False):
    """Restatement of form_factors.py:11-72 on an :class:`OracleShapeModel`."""
    P = np.ascontiguousarray(shape_model.P)
    N = np.ascontiguousarray(shape_model.N)
    A = np.ascontiguousarray(shape_model.A)
    sfx = _sfx(shape_model.dtype)
    assert P.dtype == N.dtype == A.dtype == shape_model.dtype
    if eps is None:
        eps = DEFAULT_EPS
    nf = P.shape[0]
    I = np.arange(nf, dtype=np.uint64) if I is None else np.ascontiguousarray(np.asarray(I).astype(np.uint64))
    J = np.arange(nf, dtype=np.uint64) if J is None else np.ascontiguousarray(np.asarray(J).astype(np.uint64))
    m, n = len(I), len(J)
    L = lib()
    h = getattr(L, 'oracle_ff_assemble_' + sfx)(
        shape_model.scene._h, _ptr(P), _ptr(N), _ptr(A), _ptr(I), m, _ptr(J), n,
        float(eps), int(shape_model.use_bvh), shape_model.nthreads)
    try:
        nnz = L.oracle_ff_nnz(h)
        tested = L.oracle_ff_tested(h)
        indptr = np.empty(m + 1, np.int64)
        indices = np.empty(nnz, np.int64)
        data = np.empty(nnz, shape_model.dtype)
        L.oracle_ff_copy(h, _ptr(indptr), _ptr(indices), _ptr(data))
    finally:
        L.oracle_ff_free(h)
    FF = scipy.sparse.csr_matrix((data, indices, indptr), shape=(m, n))
    if return_stats:
        return FF, {'pairs_all': m * n, 'pairs_tested': int(tested), 'nnz': int(nnz)}
    return FF


def form_factor_dense_f64(P, N, A, I=None, J=None):
    """Un-culled, un-occluded midpoint-rule kernel (P1) in float64 NumPy --
    the ground truth the value tolerances are judged against."""
    P = np.asarray(P, np.float64)
    N = np.asarray(N, np.float64)
    A = np.asarray(A, np.float64)
    nf = P.shape[0]
    I = np.arange(nf) if I is None else np.asarray(I, np.int64)
    J = np.arange(nf) if J is None else np.asarray(J, np.int64)
    d = P[J][None, :, :] - P[I][:, None, :]
    a = np.maximum(0, np.einsum('ik,ijk->ij', N[I], d))
    b = np.maximum(0, -np.einsum('jk,ijk->ij', N[J], d))
    num = a * b
    num[I[:, None] == J[None, :]] = 0
    r2 = (d * d).sum(-1)
    with np.errstate(divide='ignore', invalid='ignore'):
        val = np.where(r2 > 0, num * A[J][None, :] / (np.pi * r2 * r2), 0.0)
    return num, val
