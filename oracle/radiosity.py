"""Radiosity / steady-state restatement used ONLY as the parity metric for the
assembled matrices (TEST INFRASTRUCTURE; the thermal model itself is out of
scope).

Follows, with an iteration cap added (the reference loop has none):

* ``_solve_radiosity_jacobi_right``   src/flux/solve.py:36-45 (+ tol rule :6-8)
* ``compute_steady_state_temp``       src/flux/model.py:8-24 (E.ndim == 1 branch)
* ``get_direct_irradiance``           src/flux/shape.py:190-222 (single sun vector)
"""
import numpy as np

SIGMA_SB = 5.670374419e-8   # scipy.constants.Stefan_Boltzmann


def solve_radiosity_jacobi_right(FF, E, rho=1, tol=None, maxiter=10000):
    if tol is None:
        tol = np.finfo(E.dtype).resolution*abs(E).max()
    B = E.copy()
    for niter in range(1, maxiter + 1):
        B1 = E + FF@(rho*B)
        if abs(B1 - B).max() <= tol:
            return B, niter
        B = B1
    raise RuntimeError('Jacobi radiosity iteration did not converge (SURVEY P13)')


def compute_steady_state_temp(FF, E, rho, emiss, Fsurf=0.0, maxiter=10000):
    B = np.maximum(0, solve_radiosity_jacobi_right(FF, E, rho, maxiter=maxiter)[0])
    IR = FF@((1 - rho)*B + Fsurf)
    Q = np.maximum(0, solve_radiosity_jacobi_right(FF, IR, 1, maxiter=maxiter)[0])
    tot = np.maximum(0, (1 - rho)*B + emiss*Q + Fsurf)
    return (tot/(emiss*SIGMA_SB))**0.25


def direct_irradiance(shape_model, F0, Dsun):
    lit = ~shape_model.is_occluded(np.arange(shape_model.num_faces), Dsun)
    E = np.zeros(shape_model.num_faces, dtype=shape_model.dtype)
    E[lit] = F0*np.maximum(0, shape_model.N[lit]@Dsun)
    return E
