"""Row-sharded assembly over the GPUs of one box.

Rows of F are independent (reference src/flux/form_factors.py:45-70 carries no
cross-row state except the running ``indptr`` sum, :54, :70), so the mesh and
its LBVH are replicated on every GPU, each rank assembles one contiguous slab
of ``I`` and the ONLY exchange is an all-gather of the per-row counts, from
which every rank derives the global ``indptr`` and the offset of its slab.
``data`` / ``indices`` never cross GPUs.

One process per GPU, ``torch.distributed`` as plumbing (NCCL over NVLink on
the box, gloo in the CPU tests).
"""
import numpy as np

from . import _lib, config


def bind_to_gpu_cpus(device):
    """Restrict this process -- and the threads it starts afterwards: the library's index-expansion pool, the
    first touch of the page-locked output arena -- to the CPUs NVML reports as local to GPU ``device`` (its
    NUMA node).  On a multi-socket box the copy-out of a rank then lands in memory next to its GPU and does not
    cross the socket interconnect; on a single-node box (or without NVML) it changes nothing.  Returns the CPU
    list it bound to, or None."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63)//64)
        cpus = {64*w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def slab_bounds(m, world_size, weights=None):
    """``starts[world_size+1]`` of the contiguous row slabs (C ABI:
    ``fluxb200_slab_plan``); ``weights`` (e.g. row counts of an earlier pass)
    balances by work instead of by rows."""
    return _lib.slab_plan(m, world_size, weights)


def exchange_row_counts(local_counts, starts, group=None, device=None):
    """All-gather the per-row counts of every slab -> global int64 ``indptr``
    (length m+1) on every rank.  ``local_counts``: int64[rows of my slab]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = np.diff(starts).astype(np.int64)
    assert len(sizes) == world and len(local_counts) == sizes[rank]
    width = int(sizes.max()) if world else 0
    dev = device if device is not None else 'cpu'
    mine = torch.zeros(max(width, 1), dtype=torch.int64, device=dev)
    if sizes[rank]:
        mine[:sizes[rank]] = torch.as_tensor(np.asarray(local_counts, np.int64), device=dev)
    gathered = torch.empty(world*max(width, 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    g = gathered.cpu().numpy().reshape(world, max(width, 1))
    counts = np.concatenate([g[r, :sizes[r]] for r in range(world)]) if world else np.zeros(0, np.int64)
    indptr = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr


class SlabResult:
    """What one rank holds after a sharded assembly."""

    def __init__(self, row_start, row_stop, global_indptr, local_csr, stats, device_csr=None):
        self.row_start, self.row_stop = int(row_start), int(row_stop)
        self.global_indptr = global_indptr
        self.local_csr = local_csr       # scipy CSR of my rows (None in device-resident mode)
        self.device_csr = device_csr     # DeviceCsrSlab of my rows (device-resident mode)
        self.stats = stats

    @property
    def nnz_offset(self):
        return int(self.global_indptr[self.row_start])


def get_form_factor_matrix_sharded(shape_model, I=None, J=None, eps=None, group=None,
                                   to_host=True, weights=None):
    """Every rank calls this with the same arguments and its own
    ``CudaTrimeshShapeModel`` (same mesh, its own device).  Returns this rank's
    :class:`SlabResult`; the full matrix is the vertical stack of the slabs in
    rank order."""
    import scipy.sparse
    import torch.distributed as dist
    if eps is None:
        eps = config.DEFAULT_EPS
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nf = shape_model.num_faces
    I = np.arange(nf, dtype=np.int64) if I is None else np.asarray(I).astype(np.int64)
    starts = slab_bounds(len(I), world, weights)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    local = None
    if to_host:
        m, n, ip, ix, dv, counts, st = shape_model._ff_assemble_host(I[lo:hi], J, eps, want_row_counts=True)
        local = scipy.sparse.csr_matrix((dv, ix, ip), shape=(m, n), copy=False)
        local.has_sorted_indices = True
    else:
        import ctypes
        from .device_csr import DeviceCsrSlab
        m, n, counts, st = shape_model._ff_assemble_device(I[lo:hi], J, eps, 4, want_row_counts=True)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().fluxb200_ff_detach_csr(shape_model._handle, ctypes.byref(h)))
        dcsr = DeviceCsrSlab(h, shape_model.device, lo, len(I))
    dev = None
    if dist.get_backend(group) == 'nccl':
        import torch
        dev = torch.device('cuda', shape_model.device)
    # the one collective of the path: per-row counts -> global indptr on every rank
    indptr = exchange_row_counts(counts, starts, group, dev)
    return SlabResult(lo, hi, indptr, local, st.as_dict(), None if to_host else dcsr)
