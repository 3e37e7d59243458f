#!/usr/bin/env python
"""Generate tests/golden/* by running the reference's OWN Python for the hot path.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference,
which does not exist on the GPU box); its outputs are committed so the tests
never read /root/reference.

What executes here is the unmodified reference code:

* ``flux.form_factors.get_form_factor_matrix``        src/flux/form_factors.py:11-72
* ``flux.shape.EmbreeTrimeshShapeModel``               src/flux/shape.py:295-421
* ``flux.model.compute_steady_state_temp``             src/flux/model.py:8-24
* ``flux.solve.solve_radiosity`` (Jacobi, right)       src/flux/solve.py:4-45
* ``flux.shape.TrimeshShapeModel.get_direct_irradiance`` src/flux/shape.py:190-244

with two shims: ``cached_property`` (package absent; functools equivalent) and
an ``embree`` stand-in (oracle/embree_standin.py) whose ``intersect1M`` /
``occluded1M`` are the C oracle's restatement of Embree's robust closest-hit
kernels -- the one piece that cannot be run here.  ``flux.thermal`` (a Cython
extension not on the hot path) is stubbed so ``flux.model`` imports.

    python -m oracle.make_golden            # from the repo root
"""
import functools
import hashlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('FLUX_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


def import_reference():
    cp = types.ModuleType('cached_property')
    cp.cached_property = functools.cached_property
    sys.modules['cached_property'] = cp
    from oracle import embree_standin
    embree_standin.install()
    th = types.ModuleType('flux.thermal')
    th.PccThermalModel1D = None
    sys.path.insert(0, os.path.join(REF, 'src'))
    import flux
    sys.modules['flux.thermal'] = th
    import flux.shape
    import flux.form_factors
    import flux.model
    import flux.quadtree
    import flux.octree
    return flux, embree_standin


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(FF):
    """Size-independent summary of a CSR matrix (indices as int64, data as stored)."""
    FF.sort_indices()
    rs = np.asarray(FF.sum(axis=1, dtype=np.float64)).ravel()
    return {
        'shape': list(FF.shape), 'nnz': int(FF.nnz), 'dtype': FF.dtype.name,
        'sum': float(FF.data.astype(np.float64).sum()),
        'row_sum_min': float(rs.min()) if rs.size else 0.0,
        'row_sum_max': float(rs.max()) if rs.size else 0.0,
        'indptr_sha256': sha(FF.indptr.astype(np.int64)),
        'indices_sha256': sha(FF.indices.astype(np.int64)),
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    flux, standin = import_reference()
    from oracle import oracle
    from fluxpy_b200 import meshes
    Embree = flux.shape.EmbreeTrimeshShapeModel
    gffm = flux.form_factors.get_form_factor_matrix
    digests = {'numpy': np.__version__,
               'generator': 'oracle/make_golden.py (reference Python + embree stand-in)'}

    # ---- 1. the reference's own sphere fixtures (tests/common.py:1-4) -----------
    for name in ('icosa_sphere', 'icosa_sphere_5'):
        d = np.load(os.path.join(REF, 'tests', 'data', name + '.npz'))
        V, F = d['V'], d['F']
        arrays = {'V': V, 'F': F}
        for dt in (np.float64, np.float32):
            tag = np.dtype(dt).name
            sm = Embree(V.astype(dt), F)
            outward = (sm.P*sm.N).sum(1) > 0        # tests/test_form_factors.py:33-34
            sm.N[outward] *= -1
            FF = gffm(sm)
            digests[f'{name}/inward/{tag}'] = dict(digest(FF), F01=float(FF[0, 1]))
            arrays[f'inward_{tag}_data'] = FF.data
            if dt == np.float64:
                arrays['inward_indices'] = FF.indices.astype(np.int32)
                arrays['inward_indptr'] = FF.indptr.astype(np.int32)
                # un-oriented visibility of the inward sphere (tests/test_shape.py:54-60);
                # the Embree backend returns True on the diagonal (masked pair)
                vis = sm.get_visibility_matrix(oriented=False)
                digests[f'{name}/vis_offdiag_all_true'] = bool(
                    (vis | np.eye(len(F), dtype=bool)).all())
                digests[f'{name}/vis_diag_embree'] = bool(np.diag(vis).all())
            sm.N *= -1                               # tests/test_form_factors.py:48
            FFo = gffm(sm)
            digests[f'{name}/outward/{tag}'] = digest(FFo)
            # is_occluded on the outward sphere (tests/test_shape.py:62-84)
            D = np.array([0.3, -0.5, 0.81], dtype=dt)
            D /= np.linalg.norm(D)
            occ = sm.is_occluded(np.arange(sm.num_faces), D)
            # exact away from grazing incidence: a ray leaving 1e-3 above a face
            # almost tangentially can clear the polyhedron's edge (Embree too)
            clear = abs(sm.N@D) > 0.05
            assert (occ == (sm.N@D < 0))[clear].all()
            arrays[f'occ_{tag}'] = np.packbits(occ)
            arrays[f'occ_D_{tag}'] = D
        if name == 'icosa_sphere':
            np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)
        else:   # 500 faces: keep the mesh, the f64 pattern is implied (all but diagonal)
            np.savez_compressed(os.path.join(OUT, name + '.npz'), V=V, F=F,
                                inward_float64_data=arrays['inward_float64_data'],
                                inward_float32_data=arrays['inward_float32_data'])

    # ---- 2. rough Gaussian craters: real occlusion --------------------------------
    rng = np.random.default_rng(1234)
    for n, seed, full in ((16, 0, True), (24, 1, False), (40, 2, False)):
        name = f'crater_n{n}_s{seed}'
        arrays = {}
        for dt in (np.float64, np.float32):
            tag = np.dtype(dt).name
            V, F = meshes.gaussian_crater(n, seed, dtype=dt)
            sm = Embree(V, F, meshes.upward_normals(V, F))
            nf = sm.num_faces
            standin.use_bvh = False       # definition: brute force closest hit
            FF = gffm(sm)
            org, dr = standin.last_rays['org'], standin.last_rays['dir']
            standin.use_bvh = True
            FFb = gffm(sm)
            assert (FF != FFb).nnz == 0 and np.array_equal(FF.indices, FFb.indices)
            digests[f'{name}/full/{tag}'] = digest(FF)
            vis = sm.get_visibility_matrix()
            digests[f'{name}/vis/{tag}'] = {'sha256': sha(np.packbits(vis)),
                                            'count': int(vis.sum())}
            # rectangular, non-contiguous, unsorted index sets (per-block assembly,
            # compressed_form_factors.py:556-560)
            I = rng.permutation(nf)[:nf//3].astype(np.uintp)
            J = rng.permutation(nf)[:nf//2].astype(np.uintp)
            FB = gffm(sm, I, J)
            S = FF[I, :][:, J]                          # block == slice, bit for bit
            S.sort_indices()                            # (tests/test_compressed_form_factors.py:63-69)
            assert np.array_equal(S.indices, FB.indices)
            if dt == np.float64:
                assert np.array_equal(S.data, FB.data)
            else:   # NumPy's float32 matvec rounds differently for different len(J)
                assert np.allclose(S.data, FB.data, rtol=1e-5, atol=0)
            digests[f'{name}/block/{tag}'] = digest(FB)
            arrays[f'I_{tag}'], arrays[f'J_{tag}'] = I, J
            # sun occlusion + direct irradiance + steady state (model.py:8-24)
            e0 = np.deg2rad(10.0)
            Dsun = np.array([np.cos(e0), 0, np.sin(e0)], dtype=dt)
            E = sm.get_direct_irradiance(1365.0, Dsun)
            occ = sm.is_occluded(np.arange(nf), Dsun)
            T = flux.model.compute_steady_state_temp(FF, E.astype(np.float64), 0.12, 0.95)
            arrays[f'occ_{tag}'] = np.packbits(occ)
            arrays[f'Dsun_{tag}'] = Dsun
            arrays[f'E_{tag}'] = E
            arrays[f'T_{tag}'] = T
            arrays[f'vis_{tag}'] = np.packbits(vis)
            arrays[f'vis_shape_{tag}'] = np.array(vis.shape)
            # ray set-up of the last row (shape.py:357-383) as the reference built it
            arrays[f'lastrow_org_{tag}'] = org
            arrays[f'lastrow_dir_{tag}'] = dr
            if full:
                arrays[f'data_{tag}'] = FF.data
                arrays[f'indices_{tag}'] = FF.indices.astype(np.int32)
                arrays[f'indptr_{tag}'] = FF.indptr.astype(np.int32)
                arrays[f'block_data_{tag}'] = FB.data
                arrays[f'block_indices_{tag}'] = FB.indices.astype(np.int32)
                arrays[f'block_indptr_{tag}'] = FB.indptr.astype(np.int32)
            else:
                arrays[f'rowcounts_{tag}'] = np.diff(FF.indptr).astype(np.int32)
                arrays[f'rowsums_{tag}'] = np.asarray(FF.sum(axis=1, dtype=np.float64)).ravel()
            print(name, tag, 'nnz', FF.nnz, 'of', nf*nf, 'vis', int(vis.sum()),
                  'occluded-from-sun', int(occ.sum()))
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)

    # ---- 3. quadrant / octant index sets (quadtree.py:5-18, octree.py:5-18) ---------
    V, F = meshes.gaussian_crater(16, 0, dtype=np.float64)
    P = V[F].mean(axis=1)
    np.savez_compressed(
        os.path.join(OUT, 'block_inds.npz'), P=P,
        **{f'quad{k}': I for k, I in enumerate(flux.quadtree.get_quadrant_order(P[:, :2]))},
        **{f'oct{k}': I for k, I in enumerate(flux.octree.get_octant_order(P))})

    with open(os.path.join(OUT, 'digests.json'), 'w') as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
