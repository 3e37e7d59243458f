"""Throughput of the per-block assembly as CompressedFormFactorMatrix drives it (reference
src/flux/compressed_form_factors.py:551-567: get_form_factor_matrix for every pair of quadrants / octants) on
the 50k-face crater G(159, 0): 4 x 4 quadrant blocks and 8 x 8 octant blocks, host SciPy CSRs.

    python tools/bench_blocks.py            -> one JSON line

Compared: (a) this round's path -- face arrays compared on the host and re-sent only when they changed, the
handle's cache of prepared column sets (4 / 8 sorts + gathers instead of 16 / 64); (b) the round-1 behaviour --
P, N, A uploaded and the columns sorted and gathered on every call (`colset_cache` 0, the host-side comparison
defeated); (c) one call for the full matrix, as the lower bound.  All results are compared entry for entry.
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import blocks, meshes


def run(sm, parts, resend=False):
    t = time.perf_counter()
    out = []
    for I in parts:
        row = []
        for J in parts:
            if resend:
                sm._sent_face_data = None          # what every call did in round 1: upload P, N, A again
            row.append(fluxpy_b200.get_form_factor_matrix(sm, I, J))
        out.append(row)
    return out, time.perf_counter() - t


def main():
    V, F = meshes.gaussian_crater(159, 0, dtype=np.float32)
    N = meshes.upward_normals(V, F)
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N.copy())
    old = fluxpy_b200.CudaTrimeshShapeModel(V, F, N.copy())
    old.set_option('colset_cache', 0)
    res = {'faces': sm.num_faces}
    for name, parts in (('quadrants_4x4', blocks.get_quadrant_order(sm.P[:, :2])),
                        ('octants_8x8', blocks.get_octant_order(sm.P))):
        parts = [p for p in parts if len(p)]
        run(sm, parts), run(old, parts, True)                       # warm-up: page-locked arena, buffers
        B, t_new = run(sm, parts)
        Bo, t_old = run(old, parts, True)
        same = all((a != b).nnz == 0 and np.array_equal(a.indices, b.indices) for ra, rb in zip(B, Bo) for a, b in zip(ra, rb))
        nnz = sum(b.nnz for r in B for b in r)
        res[name] = {'blocks': len(parts)**2, 'nnz': int(nnz), 't_s': t_new, 't_round1_behaviour_s': t_old,
                     'speedup': t_old/t_new, 'identical': bool(same),
                     'colset_cache_hits': sm.trace_counters()['colset_cache_hits']}
        del B, Bo
    # one call for the whole matrix, as a user makes it: ONCE (a 10 GB result of a call shape seen for the first
    # time goes to ordinary memory through staging slots; a second call of the same shape would lock 11 GB of
    # host memory for it first, 6 s -- the rule is made for slabs in a loop, not for this)
    t = time.perf_counter()
    FF = fluxpy_b200.get_form_factor_matrix(sm)
    res['full_matrix_one_call_s'] = time.perf_counter() - t
    res['full_matrix_nnz'] = int(FF.nnz)
    res['full_matrix_pageable_results'] = int(type(sm).pageable_results)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
