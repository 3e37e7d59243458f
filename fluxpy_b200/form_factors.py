"""Mirror of ``flux.form_factors.get_form_factor_matrix``
(reference src/flux/form_factors.py:11-72) on the fused CUDA path.

One FFI call per matrix (``fluxb200_ff_assemble``: row sub-slabs traced, filled
and copied out in a two-stream pipeline into page-locked host buffers) instead
of one ray-tracing call per row; the result is the same ``scipy.sparse.csr_matrix``: shape
``(len(I), len(J))``, ``data`` in ``shape_model.dtype``, column indices =
positions into ``J`` ascending within each row, index arrays int32 when they
fit (what SciPy makes of the reference's uint64 arrays) and int64 otherwise.
"""
import numpy as np
import scipy.sparse

from . import config

#: stats of the most recent assembly in this process (pairs, nnz, stage times)
last_stats = {}


def get_form_factor_matrix(shape_model, I=None, J=None, eps=None):
    if eps is None:
        eps = config.DEFAULT_EPS
    if shape_model.dtype not in (np.float32, np.float64):
        raise RuntimeError(f'unsupported dtype {shape_model.dtype}')   # form_factors.py:37
    if not hasattr(shape_model, '_ff_assemble_host'):
        raise RuntimeError('get_form_factor_matrix needs a CudaTrimeshShapeModel: '
                           'this package has no CPU path')
    m, n, indptr, indices, data, _, st = shape_model._ff_assemble_host(I, J, eps)
    last_stats.clear()
    last_stats.update(st.as_dict())
    FF = scipy.sparse.csr_matrix((data, indices, indptr), shape=(m, n), copy=False)
    FF.has_sorted_indices = True
    return FF
