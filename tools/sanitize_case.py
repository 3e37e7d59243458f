"""Small assembly + queries for compute-sanitizer (memcheck / racecheck / initcheck).

    compute-sanitizer --tool racecheck python tools/sanitize_case.py [option=value ...]   e.g. horizon_skip=1
"""
import sys
import numpy as np
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, get_form_factor_matrix_device
for n, dt in ((20, np.float32), (14, np.float64)):
    V, F = meshes.gaussian_crater(n, 1, dtype=dt)
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
    for item in sys.argv[1:]:
        name, _, value = item.partition('=')
        sm.set_option(name, int(value))
    FF = fluxpy_b200.get_form_factor_matrix(sm)
    rng = np.random.default_rng(0)
    I = rng.permutation(sm.num_faces)[:50]; J = rng.permutation(sm.num_faces)[:300]
    FB = fluxpy_b200.get_form_factor_matrix(sm, I, J)
    vis = sm.get_visibility(I, J)
    occ = sm.is_occluded(np.arange(sm.num_faces), np.array([0.5, 0.1, 0.86], dt))
    sm.intersect1(np.array([0., 0., 1.]), np.array([0., 0., -1.]))
    D = get_form_factor_matrix_device(sm)
    y = D @ np.ones(sm.num_faces)
    print(n, dt.__name__, FF.nnz, FB.nnz, int(vis.sum()), int(occ.sum()), float(y.sum()))
