"""Fuzz the library's CUDA kernels on the SIMT emulator against the CPU oracle (test infrastructure).

Random small meshes of the kinds a regular test suite does not hold -- a handful of faces, duplicated and
degenerate triangles, coplanar sheets, vertices shared or not, spikes, huge or tiny coordinates, flipped /
non-unit / zero normals -- with random row and column subsets (repeats allowed), both dtypes, the trace kernel's
horizon skip off and on at a random zone size.  Every CSR is compared bit for bit with the oracle's; the
emulator itself reports deadlocked collectives and out-of-bounds writes.

    python tools/simt/fuzz.py [seconds] [first_seed]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def random_mesh(rng):
    kind = rng.integers(0, 8)
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    scale = float(rng.choice([1.0, 1.0, 1e-2, 37.0, 4.0e3]))
    if kind == 0:      # tiny soup
        nt = int(rng.integers(1, 40))
        c = rng.uniform(-1, 1, (nt, 1, 3))
        V = (c + rng.normal(scale=rng.choice([0.02, 0.2, 0.8]), size=(nt, 3, 3))).reshape(-1, 3)
        F = np.arange(3*nt).reshape(nt, 3)
    elif kind == 1:    # rough height field, random grid
        from fluxpy_b200 import meshes
        n = int(rng.integers(3, 22))
        V, F = meshes.gaussian_crater(n, int(rng.integers(0, 1000)), dtype=np.float64)
    elif kind == 2:    # two parallel sheets facing each other + duplicates of some faces
        n = int(rng.integers(2, 9))
        g = np.linspace(-1, 1, n)
        X, Y = np.meshgrid(g, g, indexing='ij')
        def sheet(z, flip):
            V = np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], 1)
            idx = np.arange(n*n).reshape(n, n)
            a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
            F = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)])
            return V, (F[:, ::-1] if flip else F)
        V0, F0 = sheet(0.0, False)
        V1, F1 = sheet(float(rng.uniform(0.05, 1.0)), True)
        V = np.concatenate([V0, V1])
        F = np.concatenate([F0, F1 + len(V0)])
        dup = rng.integers(0, len(F), int(rng.integers(0, 6)))
        F = np.concatenate([F, F[dup]])                       # coincident faces: closest-hit tie rule
    elif kind == 3:    # degenerate triangles mixed into a soup
        nt = int(rng.integers(4, 60))
        V = rng.uniform(-1, 1, (3*nt, 3))
        F = np.arange(3*nt).reshape(nt, 3)
        for k in rng.integers(0, nt, int(rng.integers(1, 5))):
            m = rng.integers(0, 3)
            if m == 0:
                V[3*k + 1] = V[3*k]                           # zero-length edge
            elif m == 1:
                V[3*k + 2] = 0.5*(V[3*k] + V[3*k + 1])        # collinear
            else:
                F[k] = F[k][[0, 0, 1]]                        # repeated vertex index
    elif kind == 4:    # closed small body
        from fluxpy_b200 import meshes
        V, F = meshes.cratered_body(subdiv=int(rng.integers(0, 3)), ncraters=int(rng.integers(0, 8)),
                                    seed=int(rng.integers(0, 100)), dtype=np.float64)
    elif kind == 5:    # fan around one shared vertex + spikes
        nt = int(rng.integers(3, 50))
        ang = np.sort(rng.uniform(0, 2*np.pi, nt + 1))
        rim = np.stack([np.cos(ang), np.sin(ang), rng.normal(scale=0.3, size=nt + 1)], 1)
        V = np.concatenate([[[0, 0, float(rng.normal(scale=0.5))]], rim])
        F = np.stack([np.zeros(nt, int), 1 + np.arange(nt), 2 + np.arange(nt)], 1)
    elif kind == 7:    # several 1024-column chunks per row: the per-unit list / common-ancestor / zone logic of K4
        from fluxpy_b200 import meshes
        n = int(rng.integers(24, 44))
        V, F = meshes.gaussian_crater(n, int(rng.integers(0, 1000)), dtype=np.float64)
        V[:, 2] *= float(rng.choice([0.3, 1.0, 2.5]))     # flatter / steeper relief: more or fewer occluded rays
    else:              # translated far from the origin (float32 resolution of the coordinates matters)
        from fluxpy_b200 import meshes
        n = int(rng.integers(3, 14))
        V, F = meshes.gaussian_crater(n, int(rng.integers(0, 1000)), dtype=np.float64)
        # offsets up to 1e3 mesh extents: float32 still resolves a triangle into > 1e3 ulps.  (At 1e5 extents --
        # 20 ulps per triangle -- seed 146 of the first version found a ray lying exactly in the plane of a DISTANT
        # triangle: the per-triangle Pluecker test of the oracle's brute force accepts it at t = -0.0 (T == 0,
        # noise-level edge functions), while any BVH culls the triangle by its box.  Not a regime either the
        # reference's float32 Embree scene or this library is meaningful in; noted in DESIGN.md section 5.)
        V = V + rng.uniform(-1, 1, 3)*float(rng.choice([10.0, 1e2, 1e3]))
    V = np.ascontiguousarray(V*scale, dtype)
    F = np.ascontiguousarray(F, np.int64)
    return V, F, dtype, int(kind)


def random_normals(rng, V, F):
    """None (the shape model's own), or user-supplied: flipped / scaled / a few zeroed."""
    mode = rng.integers(0, 4)
    if mode == 0:
        return None
    a, b, c = V[F[:, 0]].astype(np.float64), V[F[:, 1]].astype(np.float64), V[F[:, 2]].astype(np.float64)
    N = np.cross(b - a, c - a)
    ln = np.linalg.norm(N, axis=1)
    N = N/np.where(ln > 0, ln, 1.0)[:, None]
    if mode >= 1:
        N[rng.random(len(F)) < 0.3] *= -1
    if mode >= 2:
        N *= rng.uniform(0.5, 2.0, (len(F), 1))
    if mode == 3 and len(F) > 2:
        N[rng.integers(0, len(F), 2)] = 0.0
    return np.ascontiguousarray(N, V.dtype)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import build_emu
    from fluxpy_b200 import _lib
    _lib.SO_PATH = build_emu.build()
    _lib._lib = None
    import fluxpy_b200
    from oracle import oracle
    t0 = time.time()
    seed, cases, skipped, queries = seed0, 0, 0, 0
    kinds = {}
    while time.time() - t0 < seconds:
        rng = np.random.default_rng(seed)
        V, F, dtype, kind = random_mesh(rng)
        N = random_normals(rng, V, F)
        nf = len(F)
        I = J = None
        if rng.random() < 0.5 or nf > 1000:
            I = rng.integers(0, nf, int(rng.integers(0, min(nf, 160) + 3))).astype(np.int64)
        if rng.random() < 0.5:
            J = rng.integers(0, nf, int(rng.integers(0, 2*nf + 3))).astype(np.int64)
        eps = float(rng.choice([1e-5, 1e-5, 1e-7, 0.0, -1.0, 1e-2]))
        zone = int(rng.choice([1, 2, 5, 16, 64, 1023]))
        try:
            sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, None if N is None else N.copy())
            # brute force IS the oracle's definition (closest hit over all triangles); its own median-split BVH is
            # an accelerator that seed 100493 of the first version caught culling a degenerate (collinear) TARGET
            # triangle whose Pluecker t is noise -- the device path tests the target directly and agreed with
            # the brute force
            om = oracle.OracleShapeModel(V, F, N=None if N is None else N.copy(), use_bvh=False)
        except RuntimeError as e:          # e.g. the LBVH depth limit on pathological inputs: must be a clean error
            skipped += 1
            seed += 1
            continue
        opts = {}
        if rng.random() < 0.5:   # tunables must never change a result
            for name, choices in (('sub_rows', [1, 3, 7, 32, 512]), ('host_expand', [0, 1]), ('fill_rows', [-1, 0, 1, 2, 8]),
                                  ('shaft_filter', [0, 1]), ('top_nodes', [0, 0, 8, 64]), ('slab_limit', [0, 4, 1 << 30]),
                                  ('blocks_per_sm', [1, 4]), ('host_threads', [0, 1, 3])):
                if rng.random() < 0.4:
                    opts[name] = int(rng.choice(choices))
                    sm.set_option(name, opts[name])
        FO = oracle.get_form_factor_matrix(om, I, J, eps)
        FO.sort_indices()
        variant = int(rng.choice([1, 2, 2]))     # trace kernel generation (2 = warp-shared queue, the default)
        sm.set_option('trace_variant', variant)
        opts['trace_variant'] = variant
        for hor in (0, 1):
            sm.set_option('horizon_skip', hor)
            if hor:
                sm.set_option('horizon_zone', zone)
            FF = fluxpy_b200.get_form_factor_matrix(sm, I, J, eps)
            FF.sort_indices()
            ok = (FF.shape == FO.shape and FF.nnz == FO.nnz and np.array_equal(FF.indptr, FO.indptr)
                  and np.array_equal(FF.indices, FO.indices)
                  and np.array_equal(FF.data.view(np.uint8), FO.data.view(np.uint8)))
            if not ok:
                print(f'MISMATCH seed {seed} kind {kind} dtype {np.dtype(dtype).name} faces {nf} horizon {hor} zone {zone} '
                      f'eps {eps} options {opts} nnz {FF.nnz} vs oracle {FO.nnz}', flush=True)
                sys.exit(1)
        if rng.random() < 0.3 and nf <= 120:
            # query hooks on the same tree: visibility (BVH == brute force on the device == oracle), sun occlusion
            Iq = np.arange(nf)
            vis = sm._get_visibility(Iq, Iq)
            vo = om.get_visibility_matrix()
            vo[Iq, Iq] = False
            d = rng.normal(size=3)
            d = (d/np.linalg.norm(d)).astype(dtype)
            if not ((vis == sm._get_visibility(Iq, Iq, _bruteforce=True)).all() and (vis == vo).all()
                    and (sm.is_occluded(Iq, d) == om.is_occluded(Iq, d)).all()):
                print(f'QUERY MISMATCH seed {seed} kind {kind} dtype {np.dtype(dtype).name} faces {nf}', flush=True)
                sys.exit(1)
            queries += 1
        del sm
        kinds[kind] = kinds.get(kind, 0) + 1
        cases += 1
        seed += 1
    print(f'fuzz ok: {cases} cases, {queries} with the query hooks (seeds {seed0}..{seed - 1}, {skipped} rejected by the library with a clean error) '
          f'in {time.time() - t0:.0f} s; by mesh kind {dict(sorted(kinds.items()))}')


if __name__ == '__main__':
    main()
