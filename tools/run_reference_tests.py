"""Run the REFERENCE's own, unmodified unit tests against the CUDA backend (build container only).

The reference's tests iterate over its plugin list ``flux.shape.trimesh_shape_models``
(tests/test_shape.py:18, tests/test_form_factors.py:19, tests/test_compressed_form_factors.py:27).
This script imports the reference package from /root/reference/src as it is, puts
``CudaTrimeshShapeModel`` in that list (the Embree and CGAL backends raise ImportError here: neither
library is in the image) and runs the reference's test files from where they lie, twice:

  loop   the reference's own ``get_form_factor_matrix`` row loop (src/flux/form_factors.py:11-72,
         unmodified) drives the backend through ``get_visibility_1_to_N`` -- the plugin seam;
  fused  ``fluxpy_b200.integration.install()`` routes ``get_form_factor_matrix`` (and the per-block
         assembly of ``CompressedFormFactorMatrix``) to the one-call fused assembly.

Without a CUDA device the library's CUDA sources run on the SIMT emulator (tools/simt, test
infrastructure).  Nothing here is read on the GPU box: /root/reference does not exist there.

    python tools/run_reference_tests.py [loop|fused|both] [unittest name filters ...]
    FLUXB200_TEST_HORIZON=64 python tools/run_reference_tests.py     # the same with the horizon skip of K4 on

Shims, none of which touches the path under test: ``cached_property`` (package absent; functools has the
same decorator), ``np.product`` (removed in NumPy 2; compressed_form_factors.py:313 still calls it), a
class-level ``LinearOperator._xp`` (this SciPy sets it in ``LinearOperator.__init__``, which the reference's
``CompressedFormFactorMatrix`` never calls), empty
``matplotlib.pyplot`` / ``meshpy.triangle`` so that ``flux.ingersoll`` imports (``test_ingersoll_crater``
needs meshpy to mesh the crater and is reported as skipped).

Two of the reference's tests cannot pass with ANY ray-tracing backend; they are run, reported, and not
counted against the backend (KNOWN_REFERENCE_DEFECTS):

  test_max_depth_2_for_stretched_sphere   dies in the reference's own ``make_block``:
      ``assert False # this is wrong---fix`` (compressed_form_factors.py:301) on the force_max_depth path;
  test_is_occluded_for_sphere             its ground truth ``N@D < 0`` ignores the ray origin
      ``P + 1e-3*N`` (shape.py:410): a face with -1e-3/inradius < N.D < 0 starts its ray above its own
      plane and leaves the triangle before crossing it, so on a convex body the ray hits nothing.  With
      320 / 500 faces and a random D about two to four faces lie in that band in most draws.  The
      adjudication below replaces the ground truth by the exact answer for a convex polyhedron (clipping
      the ray against every face's half-space, float64, no library code) and requires the backend to
      match it on every face, while the reference's own ``N@D < 0`` differs from it on exactly the faces
      the unit test complains about.
"""
import functools
import os
import sys
import types
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
NEEDS_MESHPY = ('test_ingersoll_crater',)
KNOWN_REFERENCE_DEFECTS = {
    'test_max_depth_2_for_stretched_sphere': 'assert False in the reference (compressed_form_factors.py:301)',
    'test_is_occluded_for_sphere': "ground truth N@D<0 ignores the origin offset P+1e-3*N (adjudicated below)",
}


def prepare(mode):
    """Import the reference with the CUDA backend as its only shape model; returns the flux package."""
    sys.path.insert(0, ROOT)
    cp = types.ModuleType('cached_property')
    cp.cached_property = functools.cached_property
    sys.modules.setdefault('cached_property', cp)
    import numpy as np
    if not hasattr(np, 'product'):
        np.product = np.prod
    import scipy.sparse.linalg as spla
    if not hasattr(spla.LinearOperator, '_xp'):  # SciPy >= 1.16 sets it in LinearOperator.__init__, which the
        try:                                     # reference's subclasses never call (compressed_form_factors.py:710)
            from scipy._lib._array_api import array_namespace
            spla.LinearOperator._xp = array_namespace(np.empty(0))
        except ImportError:
            pass
    for name in ('matplotlib', 'matplotlib.pyplot', 'meshpy', 'meshpy.triangle'):
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = types.ModuleType(name)
    for p in (os.path.join(REF, 'tests'), REF, os.path.join(REF, 'src')):
        sys.path.insert(0, p)
    import torch  # noqa: F401  (device probe only)
    from fluxpy_b200 import _lib
    if not torch.cuda.is_available():
        sys.path.insert(0, os.path.join(ROOT, 'tools', 'simt'))
        import build_emu
        _lib.SO_PATH = build_emu.build()
        _lib._lib = None
    import flux.shape
    import flux.form_factors
    from fluxpy_b200 import CudaTrimeshShapeModel, integration
    zone = os.environ.get('FLUXB200_TEST_HORIZON')
    if zone:  # as tests/conftest.py: every shape model with the trace kernel's horizon skip on (value = zone size)
        plain_init = CudaTrimeshShapeModel.__init__

        def init_with_horizon(self, *args, **kwargs):
            plain_init(self, *args, **kwargs)
            self.set_option('horizon_zone', int(zone))
            self.set_option('horizon_skip', 1)
        CudaTrimeshShapeModel.__init__ = init_with_horizon
    reference_loop = flux.form_factors.get_form_factor_matrix
    if mode == 'fused':
        integration.install()
        assert flux.form_factors.get_form_factor_matrix is not reference_loop
    else:
        flux.shape.trimesh_shape_models.append(CudaTrimeshShapeModel)
    # the other backends cannot be constructed in this image
    flux.shape.trimesh_shape_models[:] = [CudaTrimeshShapeModel]
    return flux


def run(mode, filters):
    prepare(mode)
    loader = unittest.TestLoader()
    suite = unittest.TestSuite()
    for mod in ('test_shape', 'test_form_factors', 'tests.test_compressed_form_factors'):
        for group in loader.loadTestsFromName(mod):
            for case in group:
                name = case.id().split('.')[-1]
                if filters and not any(f in case.id() for f in filters):
                    continue
                if name in NEEDS_MESHPY:
                    setattr(case, name, unittest.skip('meshpy is not in this image')(getattr(case, name)))
                suite.addTest(case)
    import numpy as np
    saved = np.geterr()  # the reference's setUp calls np.seterr('raise') and leaves it on
    res = unittest.TextTestRunner(verbosity=2, stream=sys.stdout).run(suite)
    np.seterr(**saved)
    bad, known = [], []
    for case, _ in res.failures + res.errors:
        name = case.id().split('.')[-1].split(' ')[0]
        if hasattr(case, 'test_case'):  # a failed subTest
            name = case.test_case.id().split('.')[-1]
        (known if name in KNOWN_REFERENCE_DEFECTS else bad).append(name)
    print(f'[{mode}] reference tests run {res.testsRun}, failures {len(res.failures)}, errors {len(res.errors)}, '
          f'skipped {len(res.skipped)}')
    for name in sorted(set(known)):
        print(f'[{mode}]   known reference defect: {name}: {KNOWN_REFERENCE_DEFECTS[name]}')
    for name in sorted(set(bad)):
        print(f'[{mode}]   FAILED on the backend: {name}')
    ok = not bad
    if not filters or any('occluded' in f for f in filters):
        ok = adjudicate_is_occluded(mode) and ok
    if mode == 'loop' and (not filters or any('cross' in f for f in filters)):
        ok = cross_check_loop_vs_fused() and ok
    return ok


def cross_check_loop_vs_fused():
    """The reference's row loop (form_factors.py:11-72, NumPy arithmetic in V.dtype, one
    get_visibility_1_to_N call per row) against the fused one-call assembly, same shape model, on a rough
    crater with real occlusion.  BASELINE.json's criteria: identical visibility pattern, F within 1e-5
    relative in float32 and 1e-12 in float64."""
    import numpy as np
    import flux.form_factors
    import fluxpy_b200
    from fluxpy_b200 import meshes
    ok = True
    for dtype, tol in ((np.float32, 1e-5), (np.float64, 1e-12)):
        V, F = meshes.gaussian_crater(24, 1, dtype=dtype)
        N = meshes.upward_normals(V, F)
        sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, N.copy())
        nf = sm.num_faces
        I = np.arange(0, nf, 3)                # every third row, all columns; and a rectangular block
        cases = [(I, None), (np.arange(nf // 2, nf), np.arange(0, nf // 2)[::-1].copy())]
        for Isel, Jsel in cases:
            A = flux.form_factors.get_form_factor_matrix(sm, Isel, Jsel)   # the reference's loop
            B = fluxpy_b200.get_form_factor_matrix(sm, Isel, Jsel)         # fused
            A.sort_indices()
            same = A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            errA = errB = float('nan')
            if same and A.nnz:
                # extended-precision evaluation of form_factors.py:46-64 on the stored entries, from the shape model's own
                # P, N, A: the yardstick both are measured by (the float32 row loop loses digits to
                # cancellation in n.d on near-grazing pairs; the fused path rounds the float64 value once)
                rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
                ii = np.asarray(Isel)[rows]
                jj = (np.arange(nf) if Jsel is None else np.asarray(Jsel))[A.indices]
                P64, N64, A64 = (x.astype(np.longdouble) for x in (sm.P, sm.N, sm.A))  # x87 extended: 64-bit mantissa
                d = P64[jj] - P64[ii]
                num = np.maximum(0, (N64[ii] * d).sum(1)) * np.maximum(0, -(N64[jj] * d).sum(1))
                F64 = num * A64[jj] / (np.longdouble(np.pi) * (d * d).sum(1) ** 2)  # pi as the float64 constant of :62
                errA = float((np.abs(A.data.astype(np.longdouble) - F64) / F64).max())
                errB = float((np.abs(B.data.astype(np.longdouble) - F64) / F64).max())
            print(f'[loop] row loop vs fused, {np.dtype(dtype).name}, {A.shape[0]} x {A.shape[1]}: nnz {A.nnz} of '
                  f'{A.shape[0] * A.shape[1]} pairs, pattern identical: {same}, dtype {A.dtype}/{B.dtype}; entrywise '
                  f'relative error against the formula in extended precision: reference loop {errA:.2e}, fused {errB:.2e} '
                  f'(tolerance {tol:g})')
            ok = ok and same and A.dtype == B.dtype and errB <= tol
    return ok


def convex_occluded(P, N, org, D):
    """Exact any-hit for a convex polyhedron given by its faces' planes (outward N through P): the ray
    org + t D, t >= 0, meets the body iff the intersection of the half-space intervals is non-empty.
    Returns (occluded, margin) with margin = tmax - tmin (its sign is the answer, its size the clearance)."""
    import numpy as np
    a = N @ D                                   # (nf,)
    b = ((org[:, None, :] - P[None, :, :]) * N[None, :, :]).sum(2)  # (m, nf): signed height of org over plane k
    with np.errstate(divide='ignore', invalid='ignore'):
        t = -b / a[None, :]
    tmax = np.where(a[None, :] > 0, t, np.inf).min(1)
    tmin = np.maximum(np.where(a[None, :] < 0, t, -np.inf).max(1), 0.0)
    inside_parallel = np.where(a[None, :] == 0, b <= 0, True).all(1)
    margin = tmax - tmin
    return (margin >= 0) & inside_parallel, margin


def adjudicate_is_occluded(mode, trials=25):
    """tests/test_shape.py:62-84 with the exact ground truth for a convex body (see the module docstring)."""
    import numpy as np
    from fluxpy_b200 import CudaTrimeshShapeModel
    eps = 1e3 * np.finfo(np.float32).resolution
    rng = np.random.default_rng(2024)
    ok = True
    for fn in ('icosa_sphere.npz', 'icosa_sphere_5.npz'):
        z = np.load(os.path.join(REF, 'tests', 'data', fn))
        sm = CudaTrimeshShapeModel(z['V'], z['F'])
        sm.N[(sm.N * sm.P).sum(1) < 0] *= -1
        faces = np.arange(sm.num_faces)
        n_ref_wrong = n_backend_wrong = n_close = 0
        for _ in range(trials):
            D = rng.standard_normal(3)
            D /= np.linalg.norm(D)
            occ = np.asarray(sm.is_occluded(faces, D)).astype(bool)
            exact, margin = convex_occluded(sm.P.astype(np.float64), sm.N.astype(np.float64),
                                            sm.P + eps * sm.N, D)
            clear = np.abs(margin) > 1e-6       # not a ray grazing the silhouette within 1e-6
            n_close += int((~clear).sum())
            n_backend_wrong += int((occ != exact)[clear].sum())
            ref_gt = sm.N @ D < 0               # the unit test's ground truth
            n_ref_wrong += int((ref_gt != exact)[clear].sum())
            # where the unit test's ground truth is right, the backend agrees with it
            ok = ok and bool((occ == ref_gt)[clear & (ref_gt == exact)].all())
        print(f'[{mode}] is_occluded adjudication, {fn}: {trials} directions x {sm.num_faces} faces: backend != exact '
              f'convex answer on {n_backend_wrong} faces; the unit test\'s N@D<0 != exact on {n_ref_wrong} faces; '
              f'{n_close} grazing within 1e-6 left out')
        ok = ok and n_backend_wrong == 0
    return ok


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'both'
    filters = sys.argv[2:]
    if not os.path.isdir(REF):
        print('reference tree not present')
        return 0
    if mode == 'both':  # one process per mode: install() patches module globals
        import subprocess
        rc = 0
        for m in ('loop', 'fused'):
            rc |= subprocess.call([sys.executable, os.path.abspath(__file__), m] + filters)
        return rc
    return 0 if run(mode, filters) else 1


if __name__ == '__main__':
    sys.exit(main())
