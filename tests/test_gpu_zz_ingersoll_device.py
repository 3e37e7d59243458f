"""GPU tier, late file on purpose (written without GPU time: a surprise here must not hide the rest of the
tier under -x): config 1 through the DEVICE-RESIDENT solver -- form factors left in HBM
(get_form_factor_matrix_device), the two Jacobi solves of model.py:8-24 on the device -- against the analytic
shadow temperature of the Ingersoll crater (reference src/flux/ingersoll.py:19-33).  The flat plane around the
crater gives the resident matrix thousands of empty rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ingersoll_device_resident_solver():
    import fluxpy_b200
    from fluxpy_b200 import meshes, solve, get_form_factor_matrix_device
    beta, rc, e0, F0, rho, emiss = np.deg2rad(40), 0.8, np.deg2rad(15), 1000.0, 0.3, 0.99
    sigma = 5.670374419e-8
    f = (1 - np.cos(beta))/2
    T_gt = (F0*np.sin(e0)*f*(1 - rho)/(1 - rho*f)*(1 + rho*(1 - f)/emiss)/sigma)**0.25
    D = np.array([np.cos(e0), 0, np.sin(e0)])
    n = 61
    V, F = meshes.ingersoll_bowl(n, dtype=np.float64)
    sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
    E = sm.get_direct_irradiance(F0, D)
    FFd = get_form_factor_matrix_device(sm)
    FFh = fluxpy_b200.get_form_factor_matrix(sm)
    assert FFd.nnz == FFh.nnz and (np.diff(FFh.indptr) == 0).sum() > 0.2*len(F)      # the plane's empty rows
    T = solve.compute_steady_state_temp(FFd, E, rho, emiss)
    Rc = np.sqrt((sm.P[:, :2]**2).sum(1))
    shadow = (Rc < rc - 4/(n - 1)) & (E == 0)
    plane = Rc > rc + 4/(n - 1)
    assert abs(np.median(T[shadow])/T_gt - 1) < 0.005
    assert np.allclose(T[plane], ((1 - rho)*F0*np.sin(e0)/(emiss*sigma))**0.25, rtol=1e-12)
    # the same two solves with SciPy products on the host copy of the matrix
    def jacobi(FF, E, r):
        B = E.copy()
        for _ in range(1000):
            B1 = E + FF@(r*B)
            if abs(B1 - B).max() <= 1e-13*abs(E).max():
                return B1
            B = B1
        raise AssertionError('no convergence')
    B = jacobi(FFh, E, rho)
    Q = emiss*jacobi(FFh, FFh@((1 - rho)*B), 1.0) + (1 - rho)*B
    assert np.allclose(T, (Q/(emiss*sigma))**0.25, rtol=1e-9)
