"""On-disk form of row-sharded form-factor matrices.

The reference stores the assembled matrix with ``scipy.sparse.save_npz``
(examples/spherical_crater/collect_data.py:141,
examples/gerlache/make_true_form_factor_matrix.py:34).  At the sizes this path
targets one host buffer cannot hold the matrix, so every rank writes its own
slab as a ``save_npz``-compatible file plus one small manifest; each slab file
is readable on its own with ``scipy.sparse.load_npz``, and ``load_sharded``
stacks them back into the reference's matrix.

Host-side only (NumPy / SciPy); nothing here touches the device.
"""
import json
import os

import numpy as np
import scipy.sparse


def slab_path(prefix, rank):
    return f'{prefix}.slab{rank:04d}.npz'


def manifest_path(prefix):
    return f'{prefix}.manifest.json'


def save_slab(prefix, rank, world, local_csr, row_start, global_shape, global_indptr=None,
              compressed=False):
    """Write this rank's rows (``scipy.sparse`` CSR) as ``<prefix>.slabRRRR.npz``;
    rank 0 also writes ``<prefix>.manifest.json``.  ``global_indptr`` (identical on
    every rank, from the row-count all-gather) goes into the manifest as per-slab
    nnz offsets so a reader can place any slab without opening the others."""
    local_csr = scipy.sparse.csr_matrix(local_csr)
    assert local_csr.shape[1] == global_shape[1]
    scipy.sparse.save_npz(slab_path(prefix, rank), local_csr, compressed=compressed)
    info = {'rank': int(rank), 'row_start': int(row_start), 'row_stop': int(row_start + local_csr.shape[0]),
            'nnz': int(local_csr.nnz)}
    with open(f'{prefix}.slab{rank:04d}.json', 'w') as f:
        json.dump(info, f)
    if rank == 0:
        man = {'format': 'fluxpy_b200 row-sharded csr, one scipy.sparse.save_npz file per slab',
               'shape': [int(global_shape[0]), int(global_shape[1])], 'world_size': int(world),
               'dtype': local_csr.dtype.name}
        if global_indptr is not None:
            man['nnz'] = int(global_indptr[-1])
        with open(manifest_path(prefix), 'w') as f:
            json.dump(man, f)
    return slab_path(prefix, rank)


def save_sharded_result(prefix, result, global_shape, world, rank, compressed=False):
    """``result``: a :class:`fluxpy_b200.sharded.SlabResult` (host or device-resident)."""
    csr = result.local_csr if result.local_csr is not None else result.device_csr.to_scipy()
    return save_slab(prefix, rank, world, csr, result.row_start, global_shape, result.global_indptr, compressed)


def load_manifest(prefix):
    with open(manifest_path(prefix)) as f:
        man = json.load(f)
    slabs = []
    for r in range(man['world_size']):
        with open(f'{prefix}.slab{r:04d}.json') as f:
            slabs.append(json.load(f))
    man['slabs'] = slabs
    return man


def load_slab(prefix, rank):
    return scipy.sparse.load_npz(slab_path(prefix, rank))


def load_sharded(prefix, rows=None):
    """The whole matrix (``rows=None``) or the slabs overlapping ``rows=(lo, hi)``,
    stacked in rank order -- the reference's ``scipy.sparse.load_npz`` result."""
    man = load_manifest(prefix)
    parts, covered = [], 0
    for s in man['slabs']:
        assert s['row_start'] == covered, 'slabs must tile the rows in rank order'
        covered = s['row_stop']
        if rows is not None and (s['row_stop'] <= rows[0] or s['row_start'] >= rows[1]):
            continue
        M = load_slab(prefix, s['rank'])
        assert M.shape == (s['row_stop'] - s['row_start'], man['shape'][1]) and M.nnz == s['nnz']
        if rows is not None:
            lo, hi = max(rows[0], s['row_start']) - s['row_start'], min(rows[1], s['row_stop']) - s['row_start']
            M = M[lo:hi]
        parts.append(M)
    assert covered == man['shape'][0]
    if not parts:
        return scipy.sparse.csr_matrix((0, man['shape'][1]), dtype=np.dtype(man['dtype']))
    return scipy.sparse.vstack(parts, format='csr')
