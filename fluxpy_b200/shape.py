"""Host-side mirror of ``flux.shape`` for the form-factor path.

Same names, argument meaning and error behaviour as the reference
(src/flux/shape.py): the geometry helpers (shape.py:16-45), the abstract
``TrimeshShapeModel`` (shape.py:52-258) with its four backend hooks, and the
registry list ``trimesh_shape_models`` (shape.py:424-427).  The one backend
here, ``CudaTrimeshShapeModel``, answers every hook from ``libfluxb200.so``
(LBVH with surface-fitted slabs + path-walk traversal on the B200) and adds the fused assembly hook that
``fluxpy_b200.form_factors.get_form_factor_matrix`` calls once per matrix
instead of once per row.  No Embree, no CGAL, no CPU fallback.
"""
import ctypes
import os
import sys
import time
from abc import ABC

import numpy as np

from . import _lib


def get_centroids(V, F):
    return V[F].mean(axis=1)


def get_cross_products(V, F):
    V0 = V[F[:, 0]]
    return np.cross(V[F[:, 1]] - V0, V[F[:, 2]] - V0)


def get_face_areas(V, F):
    C = get_cross_products(V, F)
    return np.sqrt(np.sum(C**2, axis=1))/2


def get_surface_normals(V, F):
    C = get_cross_products(V, F)
    return C/np.sqrt(np.sum(C**2, axis=1)).reshape(C.shape[0], 1)


def get_surface_normals_and_face_areas(V, F):
    C = get_cross_products(V, F)
    C_norms = np.sqrt(np.sum(C**2, axis=1))
    return C/C_norms.reshape(C.shape[0], 1), C_norms/2


class ShapeModel(ABC):
    pass


class TrimeshShapeModel(ShapeModel):
    """A shape model consisting of a single triangle mesh (shape.py:52-258).

    ``V[F]`` yields the faces.  ``N``, ``P``, ``A`` may be passed to override
    the computed normals / centroids / areas; all three stay public, mutable
    attributes (the reference tests flip ``N`` in place)."""

    def __init__(self, V, F, N=None, P=None, A=None):
        if type(self) == TrimeshShapeModel:
            raise RuntimeError("tried to instantiate TrimeshShapeModel directly")
        self.dtype = V.dtype
        self.V = V
        self.F = F
        if N is not None and N.shape[0] != F.shape[0]:
            raise Exception(
                'must pass same number of surface normals as faces (got ' +
                '%d faces and %d normals' % (F.shape[0], N.shape[0]))
        self._make_scene()
        P0, N0, A0 = self._face_geometry()
        self.P = P0                      # the reference recomputes P regardless (shape.py:104)
        self.N = N0 if N is None else N
        self.A = A0 if A is None else A
        assert self.P.dtype == self.dtype
        assert self.N.dtype == self.dtype
        assert self.A.dtype == self.dtype

    def __reduce__(self):
        # device state is rebuilt on unpickle, never pickled (shape.py:114-115)
        return (self.__class__, (self.V, self.F, self.N, self.P, self.A))

    def __repr__(self):
        return 'a TrimeshShapeModel with %d vertices and %d faces' % (
            self.num_verts, self.num_faces)

    @property
    def num_faces(self):
        return self.P.shape[0]

    @property
    def num_verts(self):
        return self.V.shape[0]

    def intersect1(self, x, d):
        """Trace one ray from `x` along `d`; returns ``(i, xt)`` of the hit or None."""
        return self._intersect1(x, d)

    def get_visibility(self, I, J, oriented=False):
        """m x n boolean visibility of centroid-to-centroid rays (shape.py:138-163)."""
        vis = self._get_visibility(I, J)
        if oriented:
            I_, J_ = np.where(vis)
            gi = np.asarray(I)[I_].astype(np.int64)
            gj = np.asarray(J)[J_].astype(np.int64)
            mask = ((self.P[gj] - self.P[gi])*self.N[gi]).sum(1) <= 0
            vis[I_[mask], J_[mask]] = False
        return vis

    def get_visibility_1_to_N(self, i, J, oriented=False):
        return self.get_visibility([i], J, oriented).ravel()

    def get_visibility_matrix(self, oriented=False):
        I = np.arange(self.num_faces, dtype=np.uintp)
        return self.get_visibility(I, I, oriented)

    def is_occluded(self, I, D):
        """Is the ray from each centroid in ``I`` along ``D`` blocked (shape.py:172-179)."""
        return self._is_occluded(I, D)

    def get_direct_irradiance(self, F0, Dsun, basemesh=None, eps=None):
        """Insolation for one sun vector, per-face sun vectors or a discretised
        extended source (shape.py:190-244)."""
        if basemesh is None:
            basemesh = self
        I = ~basemesh.is_occluded(np.arange(self.num_faces), Dsun)
        if Dsun.ndim == 1:
            E = np.zeros(self.num_faces, dtype=self.dtype)
            E[I] = F0*np.maximum(0, self.N[I]@Dsun)
        elif Dsun.ndim == 2 and Dsun.shape[0] == self.num_faces:
            if Dsun.shape[1] != 3:
                raise ValueError('need Dsun.shape[1] == 3 if Dsun.ndim == 2')
            E = np.zeros(self.num_faces, dtype=self.dtype)
            E[I] = F0*np.maximum(0, (self.N[I]*Dsun[I]).sum(1))
        elif Dsun.ndim == 2:
            if Dsun.shape[1] != 3:
                raise ValueError('need Dsun.shape[1] == 3 if Dsun.ndim == 2')
            E = np.where(I, self.N@Dsun.T, 0)
            E = np.mean(F0)*np.maximum(0, np.sum(E, axis=1)/Dsun.shape[0])
        else:
            raise RuntimeError('Dsun.ndim > 2 not implemented yet')
        return E


def _index_array(I):
    return np.ascontiguousarray(np.asarray(I).astype(np.int64))


def _pageable_buffer(nbytes):
    """``nbytes`` of ordinary host memory as a uint8 array; transparent huge pages are requested where the
    kernel offers them (2 MB pages: 512 times fewer first-touch page faults when the library's host threads
    fill a multi-gigabyte result)."""
    try:
        import mmap
        buf = mmap.mmap(-1, int(nbytes), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        try:
            buf.madvise(mmap.MADV_HUGEPAGE)
        except (AttributeError, OSError, ValueError):
            pass
        return np.frombuffer(buf, dtype=np.uint8)
    except (ImportError, OSError, ValueError):
        return np.empty(int(nbytes), np.uint8)


class CudaTrimeshShapeModel(TrimeshShapeModel):
    """B200 backend: the four hooks of shape.py:261-292 / 295-421 plus the fused
    assembly hook, all through the C ABI of ``libfluxb200.so``."""

    device = 0          # CUDA ordinal used by new instances (set per process / rank)

    def _make_scene(self):
        L = _lib.lib()       # raises if the extension is missing: no fallback
        self._dtype_code = _lib.dtype_code(self.dtype)
        V = np.ascontiguousarray(self.V)
        F = np.ascontiguousarray(self.F, dtype=np.int64)
        if V.ndim != 2 or V.shape[1] != 3 or F.ndim != 2 or F.shape[1] != 3:
            raise ValueError('V and F must have shape (*, 3)')
        h = ctypes.c_void_p()
        _lib.check(L.fluxb200_mesh_create(_lib.ptr(V), V.shape[0], _lib.ptr(F), F.shape[0],
                                          self._dtype_code, int(self.device), ctypes.byref(h)))
        self._handle = h
        self._nf = F.shape[0]

    def __del__(self):
        h, self._handle = getattr(self, '_handle', None), None
        if h and _lib._lib is not None:
            _lib._lib.fluxb200_mesh_destroy(h)

    def _face_geometry(self):
        nf = self._nf
        P = np.empty((nf, 3), self.dtype)
        N = np.empty((nf, 3), self.dtype)
        A = np.empty((nf,), self.dtype)
        _lib.check(_lib.lib().fluxb200_mesh_get_face_data(self._handle, _lib.ptr(P), _lib.ptr(N), _lib.ptr(A)))
        return P, N, A

    def _sync_face_data(self):
        """P, N, A are public mutable attributes: refresh the device copies
        before every query."""
        arrs = []
        for a, shape in ((self.P, (self._nf, 3)), (self.N, (self._nf, 3)), (self.A, (self._nf,))):
            a = np.ascontiguousarray(a)
            if a.dtype != self.dtype or a.shape != shape:
                raise RuntimeError(f'face array has dtype/shape {a.dtype}{a.shape}, expected {self.dtype}{shape}')
            arrs.append(a)
        # Only what differs from what the device already holds is sent: an exact comparison with the
        # copies kept from the last upload (a memcmp of 28 bytes per face, against a host-to-device
        # copy + pack kernel + stream synchronisation per query -- the per-block assembly of
        # CompressedFormFactorMatrix makes 16 / 64 calls on an unchanged shape model).
        sent = getattr(self, '_sent_face_data', None)
        send = [a if sent is None or not np.array_equal(a, b, equal_nan=False) else None
                for a, b in zip(arrs, sent or (None, None, None))]
        #: bytes of face data the last query really copied to the device (bench.py's h2d count)
        self.face_bytes_sent_last = sum(a.nbytes for a in send if a is not None)
        if any(a is not None for a in send):
            _lib.check(_lib.lib().fluxb200_mesh_set_face_data(self._handle, *[_lib.ptr(a) for a in send]))
            self._sent_face_data = tuple(a.copy() if s is not None else b
                                         for a, s, b in zip(arrs, send, sent or (None, None, None)))

    # ---- hooks ------------------------------------------------------------------
    def _intersect1(self, x, d):
        x = np.ascontiguousarray(x, np.float64)
        d = np.ascontiguousarray(d, np.float64)
        hit, face, t = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_double(0)
        xt = np.zeros(3)
        _lib.check(_lib.lib().fluxb200_intersect1(self._handle, _lib.ptr(x), _lib.ptr(d), ctypes.byref(hit),
                                                  ctypes.byref(face), ctypes.byref(t), _lib.ptr(xt)))
        if hit.value:
            return face.value, xt

    def _get_visibility(self, I, J, _bruteforce=False):
        I, J = _index_array(I), _index_array(J)
        self._sync_face_data()
        vis = np.empty((len(I), len(J)), np.uint8)
        fn = _lib.lib().fluxb200_visibility_bruteforce if _bruteforce else _lib.lib().fluxb200_visibility
        _lib.check(fn(self._handle, _lib.ptr(I), len(I), _lib.ptr(J), len(J), _lib.ptr(vis)))
        vis = vis.astype(bool)
        # a face does not see itself: the reference's tests expect False on the
        # diagonal (tests/test_shape.py:37-39, CGAL semantics); the Embree
        # backend's "masked pair => visible" default is not kept here
        if len(I) and len(J):
            vis[I[:, None] == J[None, :]] = False
        return vis

    def _is_occluded(self, I, D):
        I = _index_array(I)
        D = np.ascontiguousarray(D, dtype=self.dtype)
        if D.ndim != 1 and D.ndim != 2:
            raise ValueError('D.ndim should be 1 or 2')
        if D.shape[-1] != 3:
            raise ValueError('need D.shape[-1] == 3')
        self._sync_face_data()
        m = len(I)
        if D.ndim == 1:
            mode, nd, shape = 0, 1, (m,)
        elif D.shape[0] == m:
            mode, nd, shape = 1, m, (m,)            # one direction per face (Embree backend)
        else:
            mode, nd, shape = 2, D.shape[0], (m, D.shape[0])   # CGAL backend's 2-D product
        occ = np.empty(shape, np.uint8)
        _lib.check(_lib.lib().fluxb200_is_occluded(self._handle, _lib.ptr(I), m, _lib.ptr(D), nd, mode,
                                                   _lib.ptr(occ)))
        return occ.astype(bool)

    # ---- fused assembly hook -------------------------------------------------------
    def _ff_count(self, I, J, eps, want_row_counts=True):
        """Pass 1 (cull + trace + count) for rows ``I`` x columns ``J`` (None = all)."""
        self._sync_face_data()
        I = None if I is None else _index_array(I)
        J = None if J is None else _index_array(J)
        m = self._nf if I is None else len(I)
        n = self._nf if J is None else len(J)
        counts = np.empty(m, np.int64) if want_row_counts else None
        st = _lib.FFStats()
        _lib.check(_lib.lib().fluxb200_ff_count(self._handle, _lib.ptr(I), m, _lib.ptr(J), n, float(eps),
                                                _lib.ptr(counts), ctypes.byref(st)))
        self._keep = (I, J)
        return m, n, counts, st

    def _ff_fill_host(self, m, nnz, index_dtype):
        """Pass 2 into freshly allocated host arrays."""
        indptr = np.empty(m + 1, index_dtype)
        indices = np.empty(nnz, index_dtype)
        data = np.empty(nnz, self.dtype)
        st = _lib.FFStats()
        _lib.check(_lib.lib().fluxb200_ff_fill(self._handle, np.dtype(index_dtype).itemsize, 0,
                                               _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
                                               ctypes.byref(st)))
        return indptr, indices, data, st

    def _ff_fill_device(self, index_width=4):
        """Pass 2 into library-owned device buffers (device-resident CSR)."""
        st = _lib.FFStats()
        _lib.check(_lib.lib().fluxb200_ff_fill(self._handle, index_width, 2, None, None, None, ctypes.byref(st)))
        return st

    #: nnz / (m*n): sizes the streaming output buffers, with 50 % headroom, so that a retry is the exception.
    #: Kept per shape model AND per call shape (m, n) -- the largest seen for that shape, so that repeated
    #: calls of one shape settle on one recycled page-locked block (a decaying estimate made a 20-step loop
    #: re-lock 3.4 GB in its ninth step: 2.2 s) -- while a dense small block cannot inflate the buffers of
    #: large calls; a shape seen for the first time starts from the most recent ratio of the model.
    _fill_ratio = 0.55
    overflow_retries = 0
    pageable_results = 0
    #: host results above this size, of a call shape not seen before, go to ordinary (pageable) memory unless a
    #: recycled page-locked block fits
    pageable_above_bytes = 4 << 30
    #: results below this many bytes are copied to ordinary memory and their page-locked block goes back to the arena
    small_result_bytes = 1 << 20

    def _note_fill_ratio(self, m, n, nnz):
        if m*n:
            r = max(0.02, nnz/(m*n))
            ratios = self.__dict__.setdefault('_fill_ratio_by_shape', {})
            if len(ratios) > 4096:
                ratios.clear()
            ratios[(m, n)] = max(r, ratios.get((m, n), 0.0))
            self._fill_ratio = r

    def _ff_assemble_host(self, I, J, eps, want_row_counts=False, index_dtype=None):
        """Streaming assembly (``fluxb200_ff_assemble``) into page-locked host
        buffers from the arena; returns zero-copy NumPy views trimmed to nnz.
        Falls back to one retry with the exact size when the estimate was short."""
        self._sync_face_data()
        I = None if I is None else _index_array(I)
        J = None if J is None else _index_array(J)
        m = self._nf if I is None else len(I)
        n = self._nf if J is None else len(J)
        L = _lib.lib()
        counts = np.empty(m, np.int64) if want_row_counts else None
        st = _lib.FFStats()
        esz = np.dtype(self.dtype).itemsize
        ratios = self.__dict__.setdefault('_fill_ratio_by_shape', {})
        ratio = ratios.get((m, n), self._fill_ratio)
        # 50 % headroom (+ 12.5 % in the arena's block): the row slabs of one mesh differ by up to 1.4x in their
        # entry counts (crater floor against rim), and a short buffer costs a re-trace AND a new page-locked block
        # -- 1.9 s inside a timed loop on 2 GPUs, 3.2 s on 8 with every rank locking at once (r02g, r02h)
        cap = int(min(m*n, max(1024, ratio*1.5*m*n + 4096)))
        cap_min = int(min(cap, max(1024, ratio*1.15*m*n + 4096)))   # what a recycled block must at least hold
        # ... and what a NEW page-locked block of a repeating call shape is made for: twice the densest slab seen so
        # far.  Blocks live long and are handed round; one sized after a sparse rim slab was later too small for
        # the crater floor and had to be replaced inside a timed step (2.8 s, 2 GPUs, r02s)
        cap_new = int(min(m*n, max(cap, ratio*2.0*m*n + 4096))) if (m, n) in ratios else cap
        if cap < 2**31 <= cap_new:
            cap_new = 2**31 - 1                                     # (keep int32 indices when they fit)
        while True:
            idt = index_dtype or (np.int32 if max(cap, n, m + 1) < 2**31 else np.int64)
            isz = np.dtype(idt).itemsize
            layout = lambda c: (-(-(c*esz)//256)*256, -(-(c*esz)//256)*256 + -(-(c*isz)//256)*256)
            off_idx, off_ptr = layout(cap)
            need = off_ptr + (m + 1)*isz
            block = _lib.arena.try_take(need)
            if block is None and cap_min < cap:
                # a recycled block that is a little short of the wanted headroom beats locking a new one (2 s for
                # 5 GB): the estimate grows whenever a denser slab of this shape turns up, the block need not
                block = _lib.arena.try_take(layout(cap_min)[1] + (m + 1)*isz)
                if block is not None:
                    cap = int(min(cap, (block.nbytes - (m + 1)*isz - 512)//(esz + isz)))
                    off_idx, off_ptr = layout(cap)
                    need = off_ptr + (m + 1)*isz
            pageable = None
            if block is None and need > self.pageable_above_bytes and (m, n) not in ratios:
                # A large result of a call shape seen for the first time, and no recycled page-locked block
                # fits: locking one would cost about 0.25 s per GB, so it goes to ordinary memory and the
                # library stages the copy-out (C ABI destination 3: page faults + a memcpy, 0.2-0.3 s per GB,
                # no locked memory left behind).  Repeated shapes (row slabs in a loop, block assembly) always
                # get page-locked blocks: locked once, recycled by the arena, copied into at PCIe speed.
                pageable = _pageable_buffer(need)
                base, dest = pageable.ctypes.data, 3
                type(self).pageable_results += 1
            else:
                if block is None:
                    cap = max(cap, cap_new)
                    off_idx, off_ptr = layout(cap)
                    need = off_ptr + (m + 1)*isz
                    block = _lib.arena.take(need)
                base, dest = block.ptr, 0
            t_call = time.perf_counter()
            rc = L.fluxb200_ff_assemble(self._handle, _lib.ptr(I), m, _lib.ptr(J), n, float(eps), isz, dest,
                                        base + off_ptr, base + off_idx, base, cap,
                                        _lib.ptr(counts), ctypes.byref(st))
            if os.environ.get('FLUXB200_ARENA_LOG'):
                sys.stderr.write('[fluxb200 assemble] %dx%d ratio %.3f cap %.3f of dense (%s) -> rc %d, nnz %.3f of dense, %.1f ms\n' % (
                    m, n, ratio, cap/max(m*n, 1), 'pageable' if pageable is not None else 'block %.3f GB' % (block.nbytes/1e9),
                    rc, st.nnz/max(m*n, 1), 1e3*(time.perf_counter() - t_call)))
            if pageable is not None and rc != _lib.OVERFLOW:
                _lib.check(rc)
                nnz = int(st.nnz)
                self._note_fill_ratio(m, n, nnz)
                data = pageable[:nnz*esz].view(self.dtype)
                indices = pageable[off_idx:off_idx + nnz*isz].view(idt)
                indptr = pageable[off_ptr:off_ptr + (m + 1)*isz].view(idt)
                return m, n, indptr, indices, data, counts, st
            if rc == _lib.OVERFLOW:      # too many entries for the buffer, or for int32 offsets: st.nnz = the count
                if block is not None:
                    _lib.arena.discard(block)
                type(self).overflow_retries += 1
                cap = cap_min = cap_new = int(min(m*n, st.nnz + st.nnz//4))   # (room for the next, denser slab of this shape)
                if index_dtype is not None and np.dtype(index_dtype).itemsize == 4 and cap >= 2**31:
                    raise RuntimeError('int32 indices cannot hold this matrix')
                continue
            if rc:
                _lib.arena.discard(block)
                _lib.check(rc)
            break
        nnz = int(st.nnz)
        self._note_fill_ratio(m, n, nnz)
        data, indices, indptr = _lib.arena.arrays(
            block, [(0, nnz, self.dtype), (off_idx, nnz, idt), (off_ptr, m + 1, idt)])
        if nnz*(esz + isz) + (m + 1)*isz <= self.small_result_bytes:
            # many small per-block matrices must not each pin a page-locked block for their lifetime
            data, indices, indptr = data.copy(), indices.copy(), indptr.copy()
        return m, n, indptr, indices, data, counts, st

    def _ff_assemble_device(self, I, J, eps, index_width=4, want_row_counts=False):
        """Streaming assembly into library-owned device buffers (CSR stays in HBM)."""
        self._sync_face_data()
        I = None if I is None else _index_array(I)
        J = None if J is None else _index_array(J)
        m = self._nf if I is None else len(I)
        n = self._nf if J is None else len(J)
        L = _lib.lib()
        counts = np.empty(m, np.int64) if want_row_counts else None
        st = _lib.FFStats()
        cap = 0
        for attempt in range(3):
            rc = L.fluxb200_ff_assemble(self._handle, _lib.ptr(I), m, _lib.ptr(J), n, float(eps), index_width,
                                        2, None, None, None, cap, _lib.ptr(counts), ctypes.byref(st))
            if rc != _lib.OVERFLOW:
                break
            cap = int(st.nnz) + 1
        _lib.check(rc)
        return m, n, counts, st

    def device_csr(self):
        """(indptr, indices, data) device pointers + nnz of the device-resident CSR."""
        ip, ix, dv, nnz = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(_lib.lib().fluxb200_ff_device_csr(self._handle, ctypes.byref(ip), ctypes.byref(ix),
                                                     ctypes.byref(dv), ctypes.byref(nnz)))
        return ip.value, ix.value, dv.value, nnz.value

    def bvh_info(self):
        info = _lib.BvhInfo()
        _lib.check(_lib.lib().fluxb200_bvh_info_get(self._handle, ctypes.byref(info)))
        return info

    def bvh_export(self):
        info = self.bvh_info()
        nodes = np.zeros((info.num_nodes, 24), np.float32)
        leaf_face = np.zeros(self._nf, np.int32)
        _lib.check(_lib.lib().fluxb200_bvh_export(self._handle, _lib.ptr(nodes), _lib.ptr(leaf_face)))
        return nodes, leaf_face

    def set_option(self, name, value):
        _lib.check(_lib.lib().fluxb200_set_option(self._handle, name.encode(), int(value)))

    def trace_counters(self):
        """Counters of the last assembly's trace launches (``fluxb200_trace_counters``)."""
        out = np.zeros(8, np.int64)
        _lib.check(_lib.lib().fluxb200_trace_counters(self._handle, _lib.ptr(out)))
        return dict(rays=int(out[0]), batches=int(out[1]), batches_source_skip=int(out[2]),
                    rays_target_skip=int(out[3]), rays_resolved_afterwards=int(out[6]),
                    colset_cache_hits=int(out[7]))

    def cuda_stream(self):
        s = ctypes.c_void_p()
        _lib.check(_lib.lib().fluxb200_mesh_stream(self._handle, ctypes.byref(s)))
        return s.value


trimesh_shape_models = [
    CudaTrimeshShapeModel,
]
