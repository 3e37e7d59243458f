"""Index sets and driver of the per-block assembly inside the reference's
``CompressedFormFactorMatrix`` (src/flux/compressed_form_factors.py:550-568,
681-691): the root level calls ``get_form_factor_matrix(shape_model, I_q, J_q)``
for every pair of quadrants / octants / user parts.  Host-side and trivial; the
work is the same CUDA assembly with arbitrary ``I``, ``J``."""
import itertools as it

import numpy as np

from .form_factors import get_form_factor_matrix


def _orthant_order(X, bbox=None):
    X = np.asarray(X)
    d = X.shape[1]
    if bbox is not None:
        lo = np.array([b[0] for b in bbox])
        hi = np.array([b[1] for b in bbox])
    else:
        lo, hi = np.min(X, axis=0), np.max(X, axis=0)
    c = (lo + hi)/2
    below = X <= c            # "<=" goes to the lower side, ">" to the upper (quadtree.py:14)
    out = []
    for upper in it.product([False, True], repeat=d):
        sel = np.ones(X.shape[0], dtype=bool)
        for k, u in enumerate(upper):
            sel &= ~below[:, k] if u else below[:, k]
        out.append(np.where(sel)[0])
    return out


def get_quadrant_order(X, bbox=None):
    """Four index arrays, x-major ((<=,<=), (<=,>), (>,<=), (>,>)); quadtree.py:5-18."""
    return _orthant_order(np.asarray(X)[:, :2], bbox)


def get_octant_order(X, bbox=None):
    """Eight index arrays, x-major; octree.py:5-18."""
    return _orthant_order(np.asarray(X)[:, :3], bbox)


def assemble_blocks(shape_model, row_parts, col_parts=None, eps=None):
    """``blocks[i][j] = get_form_factor_matrix(shape_model, row_parts[i], col_parts[j])``
    as FormFactor2dTreeBlock / FormFactorPartitionBlock do at the root."""
    if col_parts is None:
        col_parts = row_parts
    return [[get_form_factor_matrix(shape_model, I, J, eps) for J in col_parts] for I in row_parts]
