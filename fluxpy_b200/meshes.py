"""Deterministic synthetic triangle meshes for the workloads BASELINE.json names.

The reference's real inputs (Haworth / Gerlache DEMs, the Ceres shape model,
the meshpy-generated Ingersoll crater: examples/haworth_crater/haworth.py,
examples/gerlache/make_mesh.py, examples/ceres/generate_mesh_from_topo.py,
src/flux/ingersoll.py:90-455) are downloads or need packages that are absent
offline, so every config gets a seeded stand-in of the same size and character
(SURVEY.md section 8d).  Pure NumPy, no device code.
"""
import numpy as np

#: grid sizes of the Gaussian-crater sweep -> 10 082 / 49 928 / 199 712 / 500 000 faces
GAUSSIAN_CRATER_SIZES = {'10k': 72, '50k': 159, '200k': 317, '500k': 501}


def grid_faces(n):
    """Two triangles (a, b, d), (a, d, c) per cell of an n x n vertex grid."""
    iy, ix = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing='ij')
    a = (iy*n + ix).ravel()
    b, c, d = a + 1, a + n, a + n + 1
    F = np.empty((2*a.size, 3), dtype=np.int64)
    F[0::2] = np.column_stack([a, b, d])
    F[1::2] = np.column_stack([a, d, c])
    return F


def gaussian_crater(n, seed=0, dtype=np.float32, rough=True, scale=1.0, offset=(0., 0., 0.)):
    """G(n, seed): a Gaussian bowl with 16 seeded bumps/pits on ``linspace(-1,1,n)^2``.

    ``z = -0.4 exp(-(x^2+y^2)/(2*0.35^2)) + sum_k a_k exp(-|xy-c_k|^2/(2 s_k^2))``.
    Returns ``V (n*n, 3)`` in ``dtype`` and ``F (2(n-1)^2, 3)`` int64; all face
    normals have ``N_z > 0`` (the orientation examples/spherical_crater/
    collect_data.py:114-115 enforces).
    """
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (16, 2))
    a = rng.uniform(-0.05, 0.05, 16)
    s = rng.uniform(0.03, 0.10, 16)
    g = np.linspace(-1, 1, n)
    x, y = np.meshgrid(g, g, indexing='xy')
    z = -0.4*np.exp(-(x**2 + y**2)/(2*0.35**2))
    if rough:
        for k in range(16):
            z += a[k]*np.exp(-((x - c[k, 0])**2 + (y - c[k, 1])**2)/(2*s[k]**2))
    V = np.column_stack([x.ravel(), y.ravel(), z.ravel()])*scale + np.asarray(offset)
    return np.ascontiguousarray(V.astype(dtype)), grid_faces(n)


def icosphere(subdiv, radius=1.0, dtype=np.float64):
    """Icosahedron subdivided ``subdiv`` times (20*4^subdiv faces), vertices on
    the sphere, faces wound so that geometric normals point outward."""
    t = (1 + 5**0.5)/2
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    V = [np.asarray(v, float)/np.linalg.norm(v) for v in V]
    for _ in range(subdiv):
        cache, F2 = {}, []

        def mid(p, q):
            key = (min(p, q), max(p, q))
            if key not in cache:
                w = V[p] + V[q]
                V.append(w/np.linalg.norm(w))
                cache[key] = len(V) - 1
            return cache[key]
        for p, q, r in F:
            pq, qr, rp = mid(p, q), mid(q, r), mid(r, p)
            F2 += [(p, pq, rp), (q, qr, pq), (r, rp, qr), (pq, qr, rp)]
        F = F2
    V = np.asarray(V)*radius
    return np.ascontiguousarray(V.astype(dtype)), np.asarray(F, dtype=np.int64)


def cratered_body(subdiv=6, radius=470.0, ncraters=200, seed=0, dtype=np.float32):
    """Closed, heavily self-occluding body (Ceres stand-in): an icosphere whose
    radius is perturbed by a seeded field of Gaussian craters.  81 920 faces at
    ``subdiv=6``; outward normals as in examples/ceres/generate_mesh_from_topo.py."""
    V, F = icosphere(subdiv, 1.0, np.float64)
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(ncraters, 3))
    c /= np.linalg.norm(c, axis=1)[:, None]
    depth = rng.uniform(0.01, 0.06, ncraters)
    width = rng.uniform(0.05, 0.25, ncraters)
    r = np.ones(V.shape[0])
    for k in range(ncraters):
        ang2 = 2 - 2*np.clip(V@c[k], -1, 1)      # squared chord length
        r -= depth[k]*np.exp(-ang2/(2*width[k]**2))
    V = V*r[:, None]*radius
    return np.ascontiguousarray(V.astype(dtype)), F


def ingersoll_bowl(n, beta_deg=40.0, rc=0.8, dtype=np.float64):
    """Spherical-cap ("Ingersoll") crater of rim radius ``rc`` and rim slope
    ``beta`` in the plane ``[-1,1]^2`` on an n x n grid (config 1 stand-in for
    src/flux/ingersoll.py's meshpy mesher).  The sphere has radius
    ``rc/sin(beta)``; the plane z=0 is exactly flat, so plane faces cull to zero."""
    beta = np.deg2rad(beta_deg)
    R = rc/np.sin(beta)
    zc = R*np.cos(beta)
    g = np.linspace(-1, 1, n)
    x, y = np.meshgrid(g, g, indexing='xy')
    rr2 = x**2 + y**2
    z = np.where(rr2 < rc**2, zc - np.sqrt(np.maximum(R**2 - rr2, 0)), 0.0)
    V = np.column_stack([x.ravel(), y.ravel(), z.ravel()])
    return np.ascontiguousarray(V.astype(dtype)), grid_faces(n)


def upward_normals(V, F):
    """Face normals flipped to ``N_z > 0`` (collect_data.py:114-115)."""
    V0 = V[F[:, 0]]
    C = np.cross(V[F[:, 1]] - V0, V[F[:, 2]] - V0)
    N = C/np.sqrt(np.sum(C**2, axis=1)).reshape(-1, 1)
    N[N[:, 2] < 0] *= -1
    return N.astype(V.dtype)
