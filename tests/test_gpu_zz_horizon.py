"""GPU tier, last file on purpose: the trace kernel's horizon skip (`set_option('horizon_skip', 1)`, off by
default; csrc/horizon.cuh).  The skip is exact, so everything here is bit-for-bit: on == off == oracle.  The
whole gpu tier can also be run with the skip forced on for every shape model: FLUXB200_TEST_HORIZON=1023."""
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


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('zone', [48, 1023])
def test_horizon_skip_full_matrix_vs_oracle(mods, zone, dtype):
    V, F = mods['meshes'].gaussian_crater(40, 2, dtype=dtype)
    N = mods['meshes'].upward_normals(V, F)
    N[::5] *= -1                                   # user-flipped normals: the horizons follow the CURRENT N
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, N.copy())
    om = mods['oracle'].OracleShapeModel(V, F, N=N.copy())
    sm.set_option('horizon_skip', 0)               # (the skip is the default since round 2)
    off = mods['ff'].get_form_factor_matrix(sm)
    sm.set_option('horizon_zone', zone)
    sm.set_option('horizon_skip', 1)
    on = mods['ff'].get_form_factor_matrix(sm)
    c = sm.trace_counters()
    assert same_csr(on, off) and same_csr(on, mods['oracle'].get_form_factor_matrix(om))
    assert c['batches'] > 0 and c['batches_source_skip'] > 0 and c['rays_target_skip'] > 0
    # the normals change -> the horizons are recomputed, not reused
    sm.N[:] = -sm.N
    om.N[:] = -om.N
    assert same_csr(mods['ff'].get_form_factor_matrix(sm), mods['oracle'].get_form_factor_matrix(om))


@pytest.mark.parametrize('scale', [1.0, 25.0, 1000.0])
def test_horizon_skip_sampled_rows_at_scale(mods, scale):
    """Rows of the 50k-face mesh at three length scales (the ray offset 1e-3 and the float32 resolution are
    absolute): on == off, and at unit scale most of the work takes the skip."""
    V, F = mods['meshes'].gaussian_crater(159, 0, dtype=np.float32)
    V = (V*np.float32(scale)).astype(np.float32)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F, mods['meshes'].upward_normals(V, F))
    I = np.linspace(0, sm.num_faces - 1, 192).astype(np.int64)
    sm.set_option('horizon_skip', 0)
    off = mods['ff'].get_form_factor_matrix(sm, I)
    sm.set_option('horizon_skip', 1)
    on = mods['ff'].get_form_factor_matrix(sm, I)
    c = sm.trace_counters()
    assert same_csr(on, off)
    assert c['batches_source_skip'] > 0.5*c['batches'] and c['rays_target_skip'] > 0.5*c['rays'], c


def test_horizon_skip_closed_body_and_subsets(mods):
    V, F = mods['meshes'].cratered_body(subdiv=4, ncraters=60, seed=0, dtype=np.float32)
    sm = mods['shape'].CudaTrimeshShapeModel(V, F)
    om = mods['oracle'].OracleShapeModel(V, F)
    rng = np.random.default_rng(1)
    I = rng.permutation(sm.num_faces)[:700]
    J = rng.integers(0, sm.num_faces, 3000)        # repeats, unsorted
    sm.set_option('horizon_zone', 200)
    sm.set_option('horizon_skip', 1)
    assert same_csr(mods['ff'].get_form_factor_matrix(sm, I, J), mods['oracle'].get_form_factor_matrix(om, I, J))
