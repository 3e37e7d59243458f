"""GPU tier, other geometry classes of BASELINE.json's configs: closed body with
heavy self-occlusion (config 4), float64 planetocentric coordinates (config 3
variant), Ingersoll bowl (config 1), coincident faces (closest-hit
tie rule), random triangle soup (conservativeness of the box + slab tests).
Everything is compared bit for bit with the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    import fluxpy_b200
    from fluxpy_b200 import meshes, shape, form_factors
    from oracle import oracle
    return dict(pkg=fluxpy_b200, meshes=meshes, shape=shape, ff=form_factors, oracle=oracle)


def same_csr(A, B):
    A.sort_indices()
    B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr)
            and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data))


def both(mods, V, F, N=None):
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, None if N is None else N.copy())
    om = mods['oracle'].OracleShapeModel(V, F, N=None if N is None else N.copy())
    return sm, om


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_closed_cratered_body_small(mods, dtype):
    """Ceres stand-in at 5 120 faces, outward normals: most cull survivors are occluded."""
    V, F = mods['meshes'].cratered_body(subdiv=4, ncraters=60, seed=0, dtype=dtype)
    sm, om = both(mods, V, F)
    assert ((sm.P*sm.N).sum(1) > 0).mean() > 0.99          # outward
    FF = mods['ff'].get_form_factor_matrix(sm)
    st = dict(mods['ff'].last_stats)
    FO, so = mods['oracle'].get_form_factor_matrix(om, return_stats=True)
    assert same_csr(FF, FO) and st['pairs_tested'] == so['pairs_tested']
    assert 0 < FF.nnz < st['pairs_tested']                 # real occlusion happened
    nf = sm.num_faces
    I = np.arange(0, nf, 7)
    vis = sm._get_visibility(I, np.arange(nf))
    assert (vis == sm._get_visibility(I, np.arange(nf), _bruteforce=True)).all()
    vo = om.get_visibility(I, np.arange(nf))
    vo[np.arange(len(I)), I] = False
    assert (vis == vo).all()
    rng = np.random.default_rng(0)
    D = rng.normal(size=(5, 3)).astype(dtype)
    D /= np.linalg.norm(D, axis=1)[:, None]
    for d in D:
        assert (sm.is_occluded(np.arange(nf), d) == om.is_occluded(np.arange(nf), d)).all()
    # extended source: all directions for every face (CGAL 2-D variant, aabb.pyx:77-87)
    occ2 = sm.is_occluded(np.arange(nf), D)
    assert occ2.shape == (nf, 5)
    for k in range(5):
        assert (occ2[:, k] == om.is_occluded(np.arange(nf), D[k])).all()
    E = sm.get_direct_irradiance(1365.0, D)
    assert E.shape == (nf,) and (E >= 0).all()


def test_closed_cratered_body_82k_sampled_rows(mods):
    """Config 4 size: 81 920 faces; sampled rows against the oracle."""
    V, F = mods['meshes'].cratered_body(subdiv=6, seed=0, dtype=np.float32)
    sm, om = both(mods, V, F)
    rows = np.array([0, 1234, 40000, 81919, 60001, 7])
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, rows),
                    mods['oracle'].get_form_factor_matrix(om, rows))
    st = dict(mods['ff'].last_stats)
    assert 0 < st['nnz'] < st['pairs_tested']


def test_float64_planetocentric_coordinates(mods):
    """Gerlache-like: metres, Moon-centred (|p| ~ 1.7e6), float64 model.  The
    ray tracer works on the float32 copy of the vertices exactly as Embree's
    vertex buffer does (shape.py:319-325); numerators are evaluated directly so
    they do not cancel."""
    V, F = mods['meshes'].gaussian_crater(48, 3, dtype=np.float64, scale=20e3, offset=(0., 0., -1.7374e6))
    N = mods['meshes'].upward_normals(V, F)
    sm, om = both(mods, V, F, N)
    FF = mods['ff'].get_form_factor_matrix(sm, eps=1e-5)
    FO = mods['oracle'].get_form_factor_matrix(om, eps=1e-5)
    assert same_csr(FF, FO) and FF.dtype == np.float64 and FF.nnz > 0
    # 1e-12 against the float64 ground truth on the stored entries.  eps is dimensionful
    # (SURVEY P3): in metres it lets pairs with cos ~ 1e-9 through, whose value no float64
    # evaluation can give to 1e-12 -- the bound carries the condition number |d|/(n.d)
    num, val = mods['oracle'].form_factor_dense_f64(sm.P, sm.N, sm.A)
    D = FF.toarray()
    m = D != 0
    P64 = sm.P.astype(np.float64)
    d = P64[None, :, :] - P64[:, None, :]
    dn = np.sqrt((d*d).sum(-1))
    a = np.abs(np.einsum('ik,ijk->ij', sm.N, d))
    b = np.abs(np.einsum('jk,ijk->ij', sm.N, d))
    with np.errstate(divide='ignore', invalid='ignore'):
        cond = dn/a + dn/b
    tol = 1e-12 + 8*2.0**-53*cond
    assert (np.abs(D[m] - val[m]) <= tol[m]*np.abs(val[m])).all()
    assert (cond[m] < 1e3).mean() > 0.5        # and for most entries that IS 1e-12
    nf = sm.num_faces
    I = np.arange(0, nf, 11)
    assert (sm._get_visibility(I, np.arange(nf)) == sm._get_visibility(I, np.arange(nf), _bruteforce=True)).all()


@pytest.mark.parametrize('case', [(72, np.float32, 1.0), (159, np.float32, 1.0), (72, np.float64, 1.0)])
def test_visibility_and_csr_pattern_equal_heightfield_geometry(mods, case):
    """The CUDA path against a ground truth that is neither the oracle nor a ray tracer: the clearance of
    the centroid-to-centroid segment over the piecewise-linear height field at every grid-line / diagonal
    crossing (tests/helpers.py::heightfield_clearance).  BASELINE.json's criterion for the visibility mask --
    at least 99.99 % agreement, every disagreement a grazing ray within 1e-6 -- on sampled rows x all
    columns of the 10k- and 50k-face craters: no disagreement at all outside the 1e-6 band, and the CSR rows
    hold exactly the cull survivors that the geometry calls visible.  (Unit scale only: in km units the
    reference's absolute 1e-3 ray offset no longer clears the float32 rounding of the centroid for rays within
    ~1e-2 rad of the source plane, which then hit their own triangle -- DESIGN.md section 5.)"""
    from tests import helpers
    n, dtype, scale = case
    V, F = mods['meshes'].gaussian_crater(n, 0, dtype=dtype, scale=scale)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    nf = sm.num_faces
    rows = np.linspace(0, nf - 1, 48 if n < 100 else 12).astype(np.int64)   # (the NumPy check is the slow part)
    vis = sm.get_visibility(rows, np.arange(nf))
    FF = mods['ff'].get_form_factor_matrix(sm, rows)
    stored = FF.toarray() != 0
    P = sm.P.astype(np.float64)
    checked = grazing = 0
    for r, i in enumerate(rows):
        gmin, gmax = helpers.heightfield_clearance(V, n, P[i], P, ray_offset=1e-3)
        geo = (gmin > 0) | (gmax < 0)
        clear = np.minimum(np.abs(gmin), np.abs(gmax)) > 1e-6*scale
        clear[i] = False
        assert ((vis[r] == geo) | ~clear).all()
        # a stored entry is a visible pair; a visible pair facing both ways is stored (cull: form_factors.py:46-52)
        assert not (stored[r] & ~geo & clear).any()
        d = P - P[i]
        num = np.maximum(0, d@sm.N[i].astype(np.float64))*np.maximum(0, -(d*sm.N.astype(np.float64)).sum(1))
        assert stored[r][geo & clear & (num > 1e-4*scale*scale)].all()
        checked += int(clear.sum())
        grazing += int((~clear).sum()) - 1
    assert checked > 0.99*(len(rows)*(nf - 1)) and grazing < 1e-2*checked


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_sun_occlusion_equals_heightfield_geometry(mods, dtype):
    """Next row N1 on the CUDA path against the geometric ground truth (tests/helpers.py::
    heightfield_sun_clearance): low sun over the 10k-face crater, one direction for all faces and per-face
    directions (shape.py:400-421 accepts both); get_direct_irradiance follows (shape.py:190-244)."""
    from tests import helpers
    n = 72
    V, F = mods['meshes'].gaussian_crater(n, 0, dtype=dtype)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    faces = np.arange(sm.num_faces)
    for elev, az in ((3.0, 0.3), (12.0, 2.1)):
        e = np.deg2rad(elev)
        D = np.array([np.cos(e)*np.cos(az), np.cos(e)*np.sin(az), np.sin(e)]).astype(dtype)
        gmin = helpers.heightfield_sun_clearance(V, n, sm.P, sm.N, D)
        clear = np.abs(gmin) > 1e-6
        occ = sm.is_occluded(faces, D)
        assert (occ == (gmin < 0))[clear].all() and clear.mean() > 0.995
        assert (sm.is_occluded(faces, np.tile(D, (sm.num_faces, 1))) == occ).all()
        E = sm.get_direct_irradiance(1365.0, D)
        lit = clear & (gmin > 0)
        assert np.allclose(E[lit], 1365.0*np.maximum(0, sm.N[lit]@D), rtol=1e-6) and (E[clear & (gmin < 0)] == 0).all()
        assert 0.02 < occ.mean() < 0.98


def test_ingersoll_bowl(mods):
    """Config 1 stand-in: exactly flat plane faces cull to nothing, faces inside
    the spherical cap see each other (concave), block == slice."""
    V, F = mods['meshes'].ingersoll_bowl(41, dtype=np.float64)
    N = mods['meshes'].upward_normals(V, F)
    sm, om = both(mods, V, F, N)
    FF = mods['ff'].get_form_factor_matrix(sm)
    assert same_csr(FF, mods['oracle'].get_form_factor_matrix(om))
    flat = np.abs(V[F][:, :, 2]).max(1) == 0
    rc = np.diff(FF.indptr)
    assert flat.any() and (rc[flat] == 0).all()
    inner = (np.linalg.norm(sm.P[:, :2], axis=1) < 0.6)
    sub = FF[inner, :][:, inner].toarray()
    Pin = sm.P[inner]
    far = np.linalg.norm(Pin[:, None] - Pin[None], axis=2) > 0.3   # near neighbours fall under eps
    assert (sub != 0)[far].all()
    # the spherical-cap identity: inside a sphere the point kernel is 1/(4 pi R^2)
    R = 0.8/np.sin(np.deg2rad(40.0))
    i, j = np.where(inner)[0][[3, -5]]
    assert abs(FF[i, j]/sm.A[j] - 1/(4*np.pi*R*R)) < 0.05/(4*np.pi*R*R)


def test_ingersoll_analytic_flux_and_temperature(mods):
    """Config 1 end to end against the ANALYTIC solution of the spherical-cap crater (Ingersoll et al.; reference
    src/flux/ingersoll.py:19-33 for T in shadow, examples/spherical_crater/collect_data.py:180-238 for the
    absorbed flux Q on the plane / in shadow / in the sun): beta = 40 deg, rc = 0.8, sun at 15 deg, F0 = 1000,
    rho = 0.3, emiss = 0.99 (collect_data.py:8-30).  Form factors and sun occlusion from the CUDA path, the two
    Jacobi solves of model.py:8-24 written out with SciPy products (the device-resident solver on the same
    crater: tests/test_gpu_zz_ingersoll_device.py).  The error is the mesh's: it shrinks as the grid is refined."""
    beta, rc, e0, F0, rho, emiss = np.deg2rad(40), 0.8, np.deg2rad(15), 1000.0, 0.3, 0.99
    sigma = 5.670374419e-8
    f = (1 - np.cos(beta))/2
    b = f*(emiss + rho*(1 - f))/(1 - rho*f)
    T_gt = (F0*np.sin(e0)*f*(1 - rho)/(1 - rho*f)*(1 + rho*(1 - f)/emiss)/sigma)**0.25
    D = np.array([np.cos(e0), 0, np.sin(e0)])

    def jacobi(FF, E, r):                        # solve.py:36-45, albedo on the right
        B = E.copy()
        for _ in range(1000):
            B1 = E + FF@(r*B)
            if abs(B1 - B).max() <= 1e-13*abs(E).max():
                return B1
            B = B1
        raise AssertionError('no convergence')

    med = {}
    for n in (41, 61):
        V, F = mods['meshes'].ingersoll_bowl(n, dtype=np.float64)
        sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
        FF = mods['ff'].get_form_factor_matrix(sm)
        E = sm.get_direct_irradiance(F0, D)
        B = jacobi(FF, E, rho)
        Q = emiss*jacobi(FF, FF@((1 - rho)*B), 1.0) + (1 - rho)*B
        Rc, h = np.sqrt((sm.P[:, :2]**2).sum(1)), 2/(n - 1)
        plane, crater = Rc > rc + 2*h, Rc < rc - 2*h          # leave out the cells the rim passes through
        shadow, sun = crater & (E == 0), crater & (E > 0)
        elev = np.maximum(0, np.pi/2 - np.arccos(np.clip(sm.N@D, -1, 1)))
        assert plane.sum() > 0.2*len(F) and shadow.sum() > 0.1*len(F) and sun.sum() > 0.1*len(F)
        assert np.allclose(Q[plane], (1 - rho)*F0*np.sin(e0), rtol=1e-12)
        rs = abs(Q[shadow]/((1 - rho)*F0*b*np.sin(e0)) - 1)
        ru = abs(Q[sun]/((1 - rho)*F0*(np.sin(elev[sun]) + b*np.sin(e0))) - 1)
        T = (Q/(emiss*sigma))**0.25
        med[n] = (np.median(rs), np.median(ru), abs(np.median(T[shadow])/T_gt - 1))
        assert med[n][0] < 0.03 and rs.max() < 0.10 and med[n][1] < 0.01 and ru.max() < 0.02 and med[n][2] < 0.01
    assert all(med[61][k] < med[41][k] for k in range(3))


def test_coincident_faces_tie_rule(mods):
    """Two copies of the same triangle have the same hit distance: the oracle's
    index-ordered closest hit keeps the later one.  The CUDA any-hit form must
    agree (t_k == t_j and k > j occludes j)."""
    V, F = mods['meshes'].gaussian_crater(14, 5, dtype=np.float32)
    dup = np.arange(40, 80)
    F2 = np.vstack([F, F[dup]])                      # faces 40..79 exist twice
    N = mods['meshes'].upward_normals(V, F2)
    sm, om = both(mods, V, F2, N)
    FF = mods['ff'].get_form_factor_matrix(sm)
    assert same_csr(FF, mods['oracle'].get_form_factor_matrix(om))
    nf0 = len(F)
    D = FF.toarray()
    # the earlier copy is hidden behind the later copy for every source that sees the pair
    seen_late = (D[:, nf0:nf0 + len(dup)] != 0)
    seen_early = (D[:, dup] != 0)
    assert seen_late.any() and not (seen_early & seen_late).any()


def test_random_triangle_soup(mods):
    """No surface structure at all: exercises the conservativeness of the padded
    boxes and fitted slabs (BVH == brute force == oracle)."""
    rng = np.random.default_rng(11)
    nt = 1500
    c = rng.uniform(-1, 1, (nt, 1, 3))
    V = (c + rng.normal(scale=0.08, size=(nt, 3, 3))).reshape(-1, 3).astype(np.float32)
    F = np.arange(3*nt).reshape(nt, 3)
    sm, om = both(mods, V, F)
    I = np.arange(nt)
    vis = sm._get_visibility(I, I)
    assert (vis == sm._get_visibility(I, I, _bruteforce=True)).all()
    vo = om.get_visibility_matrix()
    vo[I, I] = False
    assert (vis == vo).all() and 0.05 < vis.mean() < 0.95
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, eps=1e-7),
                    mods['oracle'].get_form_factor_matrix(om, eps=1e-7))
    d = np.array([0.2, -0.3, 0.9], np.float32)
    assert (sm.is_occluded(I, d) == om.is_occluded(I, d)).all()
    hit = sm.intersect1(np.array([0., 0., 5.]), np.array([0., 0., -1.]))
    assert hit is None or (0 <= hit[0] < nt)


def test_full_50k_matrix_properties(mods):
    """BASELINE config 2 size (G(159,0), 49 928 faces): the whole 2.5e9-pair
    matrix, device-resident; counts are consistent with a second, differently
    split assembly; sampled rows equal the oracle."""
    V, F = mods['meshes'].gaussian_crater(159, 0, dtype=np.float32)
    N = mods['meshes'].upward_normals(V, F)
    sm, om = both(mods, V, F, N)
    nf = sm.num_faces
    m, n, counts, st = sm._ff_assemble_device(None, None, 1e-5, 4, want_row_counts=True)
    assert m == n == nf and counts.sum() == st.nnz and st.pairs_all == nf*nf
    assert 0.3*nf*nf < st.nnz <= st.pairs_tested
    sm.set_option('sub_rows', 3000)
    m2, n2, counts2, st2 = sm._ff_assemble_device(None, None, 1e-5, 4, want_row_counts=True)
    assert np.array_equal(counts, counts2) and st2.pairs_tested == st.pairs_tested
    rows = np.array([5, 25000, 49927])
    FO = mods['oracle'].get_form_factor_matrix(om, rows)
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, rows), FO)
    assert np.array_equal(np.diff(FO.indptr), counts[rows])


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_device_resident_radiosity_and_steady_state(mods, dtype):
    """Next row N2: Jacobi radiosity + steady-state temperature on the
    device-resident matrix == the reference's own T (golden, model.py:8-24)
    to 1e-6 relative L2, same iteration counts as the CPU restatement."""
    from fluxpy_b200 import solve, get_form_factor_matrix_device
    from oracle import radiosity
    from tests import helpers
    tag = np.dtype(dtype).name
    g = helpers.load('crater_n24_s1')
    V, F = mods['meshes'].gaussian_crater(24, 1, dtype=dtype)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    FFd = get_form_factor_matrix_device(sm)
    FFh = mods['ff'].get_form_factor_matrix(sm)
    D = FFd.to_scipy()
    assert same_csr(D, FFh) and FFd.nnz == FFh.nnz and FFd.shape == FFh.shape
    x = np.random.default_rng(0).normal(size=FFh.shape[1])
    assert np.allclose(FFd@x, FFh@x, rtol=1e-12, atol=1e-14)
    E = sm.get_direct_irradiance(1365.0, g[f'Dsun_{tag}']).astype(np.float64)
    for rho in (0.12, np.linspace(0.05, 0.3, FFh.shape[0])):
        B, nit = solve.solve_radiosity(FFd, E, rho)
        Bref, nref = radiosity.solve_radiosity_jacobi_right(FFh, E, rho)
        assert nit == nref and np.allclose(B, Bref, rtol=1e-13, atol=1e-10)
        Bl, _ = solve.solve_radiosity(FFd, E, rho, albedo_placement='left')
        assert np.allclose(Bl[E > 0], B[E > 0], rtol=0.5) and np.isfinite(Bl).all()
    T = solve.compute_steady_state_temp(FFd, E, 0.12, 0.95)
    Tref = g[f'T_{tag}']
    assert np.linalg.norm(T - Tref) <= 1e-6*np.linalg.norm(Tref)
    with pytest.raises(RuntimeError):            # rho = 5: diverges, the reference would hang (P13)
        solve.solve_radiosity(FFd, E, 5.0, maxiter=200)
    # the mesh handle is free for the next assembly after the detach
    assert same_csr(mods['ff'].get_form_factor_matrix(sm), FFh)


def test_device_resident_slab_to_disk_roundtrip(mods, tmp_path):
    """Next row N4: a slab that was assembled device-resident (the 8-GPU mode) goes to disk as a
    ``scipy.sparse.save_npz`` file + manifest and loads back as the matrix the reference stores
    (examples/gerlache/make_true_form_factor_matrix.py:34).  One rank, gloo, so it runs in the 1-GPU tier; the
    two-rank version is in test_gpu_multi.py."""
    import scipy.sparse
    import torch.distributed as dist
    from fluxpy_b200 import sharded, io as ffio
    V, F = mods['meshes'].gaussian_crater(24, 1, dtype=np.float32)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    full = mods['ff'].get_form_factor_matrix(sm)
    mine = not dist.is_initialized()
    if mine:
        dist.init_process_group('gloo', init_method=f'file://{tmp_path}/rendezvous', rank=0, world_size=1)
    try:
        res = sharded.get_form_factor_matrix_sharded(sm, to_host=False)
        assert res.local_csr is None and res.device_csr.nnz == full.nnz
        prefix = str(tmp_path / 'ff')
        ffio.save_sharded_result(prefix, res, full.shape, 1, 0)
    finally:
        if mine:
            dist.destroy_process_group()
    back = ffio.load_sharded(prefix)
    assert same_csr(back, full) and back.dtype == full.dtype
    assert same_csr(scipy.sparse.load_npz(ffio.slab_path(prefix, 0)), full)     # a plain save_npz file
    assert ffio.load_manifest(prefix)['nnz'] == full.nnz


@pytest.mark.parametrize('scale', [1.0, 25.0, 1000.0])
def test_culling_structures_are_conservative_at_scale(mods, scale):
    """Every culling device of the trace kernel -- fitted slabs, the per-unit
    record list with its shaft filter, the upward target-path walk -- may only
    skip triangles the exact test would reject.  On the full 2.5e9-pair matrix of
    G(159,0), at three length scales (eps and the 1e-3 ray offset are
    dimensionful), the row counts must not depend on any of them.  (A first
    version of the shaft filter lost 3 occluders out of 1.3e9 rays at scale 25:
    rays grazing their target run on past its centroid; rows 13454, 23534, 26337.)"""
    V, F = mods['meshes'].gaussian_crater(159, 0, dtype=np.float32)
    V = V*np.float32(scale)
    N = mods['meshes'].upward_normals(V, F)
    sm, om = both(mods, V, F, N)
    ref = None
    for name, val in ((None, None), ('shaft_filter', 0), ('slab_limit', 0), ('top_nodes', 64)):
        if name:
            sm.set_option(name, val)
        m, n, counts, st = sm._ff_assemble_device(None, None, 1e-5, 4, want_row_counts=True)
        if ref is None:
            ref = counts
            assert st.nnz < st.pairs_tested
        assert np.array_equal(counts, ref), name
        if name == 'shaft_filter':
            sm.set_option(name, 1)
    rows = np.array([13454, 23534, 26337, 101])
    FO = mods['oracle'].get_form_factor_matrix(om, rows)
    assert np.array_equal(np.diff(FO.indptr), ref[rows])


def test_large_index_subsets_and_block_assembly(mods):
    """Per-block assembly at scale (compressed_form_factors.py:550-568): many
    chunks of non-contiguous, unsorted columns against the oracle, and the 4 x 4
    quadrant blocks of a 10k-face crater == slices of the full matrix, bit for
    bit (the property tests/test_compressed_form_factors.py:63-69 pins)."""
    from fluxpy_b200 import blocks
    V, F = mods['meshes'].gaussian_crater(159, 0, dtype=np.float32)
    N = mods['meshes'].upward_normals(V, F)
    sm, om = both(mods, V, F, N)
    rng = np.random.default_rng(7)
    nf = sm.num_faces
    I = rng.permutation(nf)[:160]
    J = rng.permutation(nf)[:30000]                       # ~30 chunks, leaf positions with gaps
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, I, J),
                    mods['oracle'].get_form_factor_matrix(om, I, J))
    Jd = np.r_[J[:5000], J[:5000]]                         # repeated columns
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, I[:20], Jd),
                    mods['oracle'].get_form_factor_matrix(om, I[:20], Jd))
    V, F = mods['meshes'].gaussian_crater(72, 0, dtype=np.float64)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    full = mods['ff'].get_form_factor_matrix(sm)
    parts = blocks.get_quadrant_order(sm.P[:, :2])
    assert sum(len(p) for p in parts) == sm.num_faces
    B = blocks.assemble_blocks(sm, parts)
    for i, Iq in enumerate(parts):
        for j, Jq in enumerate(parts):
            S = full[Iq, :][:, Jq]
            S.sort_indices()
            assert B[i][j].shape == (len(Iq), len(Jq))
            assert np.array_equal(S.indices, B[i][j].indices) and np.array_equal(S.data, B[i][j].data)


def test_block_extraction_and_lowrank_feed(mods):
    """Next row N3: device-side spmat[row_inds, :][:, col_inds]
    (compressed_form_factors.py:562) == SciPy slicing bit for bit; thin products
    A@X, A.T@X == SciPy; randomised truncated SVD of an off-diagonal quadrant
    block == scipy.sparse.linalg.svds (what linalg.py:8-18 calls)."""
    import scipy.sparse.linalg
    import torch
    from fluxpy_b200 import blocks, lowrank, get_form_factor_matrix_device
    V, F = mods['meshes'].gaussian_crater(40, 2, dtype=np.float32)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    FFd = get_form_factor_matrix_device(sm)
    FFh = FFd.to_scipy()
    parts = blocks.get_quadrant_order(sm.P[:, :2])
    for (i, j) in ((0, 3), (1, 1), (2, 0)):
        B = FFd.extract(parts[i], parts[j])
        S = FFh[parts[i], :][:, parts[j]]
        S.sort_indices()
        Bh = B.to_scipy()
        assert Bh.shape == S.shape and np.array_equal(Bh.indptr, S.indptr)
        assert np.array_equal(Bh.indices, S.indices) and np.array_equal(Bh.data, S.data)
    # unsorted / repeated rows, unsorted columns: same content as SciPy after sorting
    rng = np.random.default_rng(2)
    rows = rng.integers(0, FFh.shape[0], 200)
    cols = rng.permutation(FFh.shape[1])[:700]
    Bh = FFd.extract(rows, cols).to_scipy()
    S = FFh[rows, :][:, cols]
    assert (Bh != S).nnz == 0 and Bh.nnz == S.nnz
    empty = FFd.extract([], cols)
    assert empty.shape == (0, 700) and empty.nnz == 0
    with pytest.raises(RuntimeError):
        FFd.extract([0], [3, 3])
    # thin products
    B = FFd.extract(parts[0], parts[3])
    Bh = B.to_scipy().astype(np.float64)
    X = rng.normal(size=(Bh.shape[1], 40))
    Y = B.matmat(torch.as_tensor(X, device='cuda')).cpu().numpy()
    assert np.allclose(Y, Bh@X, rtol=1e-12, atol=1e-14)
    Z = rng.normal(size=(Bh.shape[0], 7))
    W = B.rmatmat(torch.as_tensor(Z, device='cuda')).cpu().numpy()
    assert np.allclose(W, Bh.T@Z, rtol=1e-11, atol=1e-13)
    # truncated SVD of the (numerically low-rank) far-field block
    k = 12
    U, Sg, Vt = lowrank.sparse_svd(B, k)
    Sref = np.sort(scipy.sparse.linalg.svds(Bh, k, return_singular_vectors=False))[::-1]
    assert np.allclose(Sg, Sref, rtol=1e-6)
    assert U.shape == (Bh.shape[0], k) and Vt.shape == (k, Bh.shape[1])
    approx = (U*Sg)@Vt
    err = np.linalg.norm(Bh.toarray() - approx, 2)
    assert err <= 1.05*np.linalg.svd(Bh.toarray(), compute_uv=False)[k] + 1e-12
    out = lowrank.estimate_rank(B, 1e-2)
    assert out is not None and 1 <= out[1].size <= min(Bh.shape) and out[1][-1] >= 1e-2*out[1][0]


def test_degenerate_triangles_follow_the_brute_force_definition(mods):
    """Collinear / repeated-vertex triangles have no well-defined plane: the per-triangle Pluecker test
    answers with rounding noise and a hit distance unrelated to where the triangle is, so no bounding volume
    is conservative for them.  The LBVH gives such leaves the whole scene as box and an open slab (and their
    near zone no horizon), so the tree still returns what the contract's brute force returns.  Found by
    tools/simt/fuzz.py (seeds 100493, 101531)."""
    rng = np.random.default_rng(5)
    V, F = mods['meshes'].gaussian_crater(14, 3, dtype=np.float32)
    V = V.astype(np.float64)
    extra_V, extra_F = [], []
    nv = len(V)
    for k in range(6):
        a, b = V[rng.integers(0, nv)] + [0, 0, 0.05], V[rng.integers(0, nv)] + [0, 0, 0.3]
        if k % 3 == 0:
            tri = [a, b, 0.5*(a + b)]                    # collinear
        elif k % 3 == 1:
            tri = [a, a, b]                              # zero-length edge
        else:
            tri = [a, b, a + 0.25*(b - a)]               # collinear, uneven
        extra_F.append([nv + len(extra_V) + q for q in range(3)])
        extra_V.extend(tri)
    V = np.concatenate([V, np.array(extra_V)]).astype(np.float32)
    F = np.concatenate([F, np.array(extra_F)])
    F = np.concatenate([F, [[F[0, 0], F[0, 0], F[0, 1]]]])          # repeated vertex index
    for dtype in (np.float32, np.float64):
        Vd = V.astype(dtype)
        N = mods['meshes'].upward_normals(Vd, F)
        N[~np.isfinite(N).all(1)] = [0, 0, 1]
        sm = mods['shape'].CudaTrimeshShapeModel(Vd, F, N.copy())
        brute = mods['oracle'].OracleShapeModel(Vd, F, N=N.copy(), use_bvh=False)
        tree = mods['oracle'].OracleShapeModel(Vd, F, N=N.copy(), use_bvh=True)
        for eps in (1e-5, -1.0):
            FF = mods['ff'].get_form_factor_matrix(sm, eps=eps)
            assert same_csr(FF, mods['oracle'].get_form_factor_matrix(brute, eps=eps))
            assert same_csr(FF, mods['oracle'].get_form_factor_matrix(tree, eps=eps))
        I = np.arange(sm.num_faces)
        vis = sm._get_visibility(I, I)
        assert (vis == sm._get_visibility(I, I, _bruteforce=True)).all()
        vo = brute.get_visibility_matrix()
        vo[I, I] = False
        assert (vis == vo).all()
