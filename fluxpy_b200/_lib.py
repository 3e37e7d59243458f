"""ctypes binding of ``libfluxb200.so`` (C ABI: ``include/fluxb200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device
is usable, every entry point of the package raises.  ``build()`` compiles the
library in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import sys
import time
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (FLUXB200_SO: another build of the same library, e.g. a tuning variant made by `make -C csrc VARIANT=...`)
SO_PATH = os.environ.get('FLUXB200_SO') or os.path.join(_HERE, 'libfluxb200.so')
CSRC = os.path.join(_HERE, 'csrc')

F32, F64 = 0, 1
ABI_VERSION = 7
OVERFLOW = 2

#: every symbol include/fluxb200.h declares
EXPORTS = (
    'fluxb200_last_error', 'fluxb200_abi_version', 'fluxb200_device_count',
    'fluxb200_mesh_create', 'fluxb200_mesh_destroy', 'fluxb200_mesh_set_face_data',
    'fluxb200_mesh_get_face_data', 'fluxb200_bvh_build', 'fluxb200_bvh_info_get',
    'fluxb200_bvh_export', 'fluxb200_ff_count', 'fluxb200_ff_fill', 'fluxb200_ff_assemble',
    'fluxb200_host_alloc', 'fluxb200_host_free', 'fluxb200_ff_device_csr', 'fluxb200_ff_detach_csr',
    'fluxb200_csr_destroy', 'fluxb200_csr_info', 'fluxb200_csr_to_host', 'fluxb200_csr_jacobi_step',
    'fluxb200_csr_extract', 'fluxb200_csr_matmat',
    'fluxb200_visibility', 'fluxb200_is_occluded', 'fluxb200_intersect1',
    'fluxb200_visibility_bruteforce', 'fluxb200_slab_plan', 'fluxb200_mesh_stream',
    'fluxb200_set_option', 'fluxb200_expand_words', 'fluxb200_expand_rows', 'fluxb200_trace_counters',
)


class FFStats(ctypes.Structure):
    _fields_ = [('pairs_all', ctypes.c_int64), ('pairs_tested', ctypes.c_int64),
                ('nnz', ctypes.c_int64), ('ms_prepare', ctypes.c_float),
                ('ms_trace', ctypes.c_float), ('ms_scan', ctypes.c_float),
                ('ms_fill', ctypes.c_float), ('ms_d2h', ctypes.c_float),
                ('trace_launches', ctypes.c_int32), ('kernel_launches', ctypes.c_int32),
                ('h2d_bytes', ctypes.c_int64), ('d2h_bytes', ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BvhInfo(ctypes.Structure):
    _fields_ = [('num_faces', ctypes.c_int64), ('num_nodes', ctypes.c_int64),
                ('num_top_nodes', ctypes.c_int32), ('max_depth', ctypes.c_int32),
                ('ms_build', ctypes.c_float), ('scene_lo', ctypes.c_float*3),
                ('scene_hi', ctypes.c_float*3)]


def build(force=False, verbose=False):
    """Compile ``libfluxb200.so`` for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(os.path.dirname(_HERE), 'include', 'fluxb200.h')]
    if (not force and os.path.exists(SO_PATH)
            and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return SO_PATH
    out = subprocess.run(['make', '-C', CSRC, 'all'], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout + out.stderr)
    if out.returncode:
        raise RuntimeError('building libfluxb200.so failed')
    return SO_PATH


_lib = None


def lib():
    """The loaded library; raises RuntimeError (never falls back) if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f'{SO_PATH} is missing: the CUDA extension has not been built '
            '(python -c "import __graft_entry__ as g; g.build()"); there is no CPU fallback')
    L = ctypes.CDLL(SO_PATH)
    vp, sz, i32, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int64
    pp = ctypes.POINTER(vp)
    L.fluxb200_last_error.restype = ctypes.c_char_p
    L.fluxb200_abi_version.restype = i32
    L.fluxb200_device_count.argtypes = [ctypes.POINTER(i32)]
    L.fluxb200_mesh_create.argtypes = [vp, sz, vp, sz, i32, i32, pp]
    L.fluxb200_mesh_destroy.argtypes = [vp]
    L.fluxb200_mesh_set_face_data.argtypes = [vp, vp, vp, vp]
    L.fluxb200_mesh_get_face_data.argtypes = [vp, vp, vp, vp]
    L.fluxb200_bvh_build.argtypes = [vp]
    L.fluxb200_bvh_info_get.argtypes = [vp, ctypes.POINTER(BvhInfo)]
    L.fluxb200_bvh_export.argtypes = [vp, vp, vp]
    L.fluxb200_ff_count.argtypes = [vp, vp, sz, vp, sz, ctypes.c_double, vp, ctypes.POINTER(FFStats)]
    L.fluxb200_ff_fill.argtypes = [vp, i32, i32, vp, vp, vp, ctypes.POINTER(FFStats)]
    L.fluxb200_ff_assemble.argtypes = [vp, vp, sz, vp, sz, ctypes.c_double, i32, i32, vp, vp, vp, i64, vp,
                                       ctypes.POINTER(FFStats)]
    L.fluxb200_host_alloc.argtypes = [sz, pp]
    L.fluxb200_host_free.argtypes = [vp]
    L.fluxb200_ff_detach_csr.argtypes = [vp, pp]
    L.fluxb200_csr_destroy.argtypes = [vp]
    L.fluxb200_csr_info.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64),
                                    ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_float)]
    L.fluxb200_csr_to_host.argtypes = [vp, vp, vp, vp]
    L.fluxb200_csr_jacobi_step.argtypes = [vp, vp, vp, ctypes.c_double, vp, vp, ctypes.POINTER(ctypes.c_double), i64]
    L.fluxb200_csr_extract.argtypes = [vp, vp, sz, vp, sz, pp]
    L.fluxb200_csr_matmat.argtypes = [vp, vp, i32, vp, i32]
    L.fluxb200_ff_device_csr.argtypes = [vp, pp, pp, pp, ctypes.POINTER(i64)]
    L.fluxb200_visibility.argtypes = [vp, vp, sz, vp, sz, vp]
    L.fluxb200_visibility_bruteforce.argtypes = [vp, vp, sz, vp, sz, vp]
    L.fluxb200_is_occluded.argtypes = [vp, vp, sz, vp, sz, i32, vp]
    L.fluxb200_intersect1.argtypes = [vp, vp, vp, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                      ctypes.POINTER(ctypes.c_double), vp]
    L.fluxb200_slab_plan.argtypes = [sz, i32, vp, vp]
    L.fluxb200_mesh_stream.argtypes = [vp, pp]
    L.fluxb200_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.fluxb200_trace_counters.argtypes = [vp, vp]
    L.fluxb200_expand_words.argtypes = [vp, sz, i32, vp, ctypes.POINTER(i64)]
    L.fluxb200_expand_rows.argtypes = [vp, sz, sz, vp, i32, vp, i32]
    for name in EXPORTS:
        if name != 'fluxb200_last_error' and name != 'fluxb200_abi_version':
            getattr(L, name).restype = i32
    if L.fluxb200_abi_version() != ABI_VERSION:
        raise RuntimeError('libfluxb200.so ABI version mismatch; rebuild it')
    _lib = L
    return L


def check(status):
    """Non-zero status -> RuntimeError(last_error) (SURVEY 8b 'Errors')."""
    if status:
        raise RuntimeError(lib().fluxb200_last_error().decode('utf-8', 'replace'))


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def dtype_code(dtype):
    if dtype == np.float32:
        return F32
    if dtype == np.float64:
        return F64
    raise RuntimeError(f'unsupported dtype {dtype}')   # form_factors.py:37


def device_count():
    n = ctypes.c_int(0)
    check(lib().fluxb200_device_count(ctypes.byref(n)))
    return n.value


def slab_plan(m, nranks, weights=None):
    """Contiguous row slabs ``starts[nranks+1]`` (host-only, no device needed)."""
    starts = np.zeros(nranks + 1, np.int64)
    w = None if weights is None else np.ascontiguousarray(weights, np.int64)
    check(lib().fluxb200_slab_plan(int(m), int(nranks), ptr(w), ptr(starts)))
    return starts


# ---------------------------------------------------------------------------
# page-locked host arena for the CSR outputs
# ---------------------------------------------------------------------------
class _Block:
    __slots__ = ('ptr', 'nbytes', 'leases')

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes, self.leases = ptr, nbytes, 0


class _Lease:
    """Exposes a slice of a pinned block to NumPy (``__array_interface__``); the
    arrays made from it keep it alive, and when the last one dies the block goes
    back to the arena."""

    def __init__(self, arena, block, offset, count, dtype):
        self._arena, self._block = arena, block
        block.leases += 1
        self.__array_interface__ = {'shape': (int(count),), 'typestr': np.dtype(dtype).str,
                                    'data': (block.ptr + offset, False), 'version': 3}

    def __del__(self):
        b = self._block
        b.leases -= 1
        if b.leases == 0:
            self._arena._give_back(b)


class PinnedArena:
    """Recycles page-locked host blocks (``fluxb200_host_alloc``) across calls:
    D2H copies into them run at full PCIe rate and asynchronously, and the
    returned CSR arrays are zero-copy views.  A block is reused only after every
    array viewing it has been garbage-collected."""

    def __init__(self, max_free_bytes=64 << 30):
        self.free = []
        self.max_free_bytes = max_free_bytes
        self.allocations = 0      # blocks page-locked so far (each costs ~0.4 s per GB: tests watch this)

    def try_take(self, nbytes):
        """A free block that fits, or None (no allocation)."""
        nbytes = int(max(nbytes, 1))
        best = None
        for b in self.free:
            if b.nbytes >= nbytes and (best is None or b.nbytes < best.nbytes):
                best = b
        if best is not None and best.nbytes <= 8*nbytes + (1 << 20):
            self.free.remove(best)
            return best
        return None

    def take(self, nbytes):
        nbytes = int(max(nbytes, 1))
        best = self.try_take(nbytes)
        if best is not None:
            return best
        p = ctypes.c_void_p()
        want = nbytes + nbytes//8 + 4096          # page-locking is slow: leave room to be reused
        t0 = time.perf_counter()
        check(lib().fluxb200_host_alloc(want, ctypes.byref(p)))
        self.allocations += 1
        if os.environ.get('FLUXB200_ARENA_LOG'):     # diagnostic: who page-locks what, and how long it takes
            sys.stderr.write('[fluxb200 arena] locked %.3f GB in %.2f s (block %d; free blocks: %s)\n' % (
                want/1e9, time.perf_counter() - t0, self.allocations, [round(b.nbytes/1e9, 3) for b in self.free]))
        return _Block(p.value, want)

    def _give_back(self, block):
        try:
            self.free.append(block)
            while sum(b.nbytes for b in self.free) > self.max_free_bytes and self.free:
                big = max(self.free, key=lambda b: b.nbytes)
                self.free.remove(big)
                lib().fluxb200_host_free(big.ptr)
        except Exception:      # interpreter shutdown
            pass

    def release_free(self):
        """Unlock and free every block that is not in use."""
        while self.free:
            lib().fluxb200_host_free(self.free.pop().ptr)

    def discard(self, block):
        """Return a block that was never leased."""
        self._give_back(block)

    def arrays(self, block, specs):
        """specs: [(offset_bytes, count, dtype)] -> NumPy arrays viewing the block."""
        return [np.asarray(_Lease(self, block, off, cnt, dt)) for off, cnt, dt in specs]


arena = PinnedArena()
