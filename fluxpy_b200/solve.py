"""Mirror of ``flux.solve.solve_radiosity`` (reference src/flux/solve.py:4-45,
Jacobi) for form-factor matrices that live on the GPU(s).

``FF`` is a :class:`~fluxpy_b200.device_csr.DeviceCsrSlab`: the whole matrix on
one GPU, or this rank's row slab of a row-sharded matrix (then every rank calls
with the same ``E``/``rho`` and the iterate is all-gathered once per iteration:
8*Nf bytes over NVLink -- the only exchange).

Same iteration, same stopping rule, same return value (the LAST-BUT-ONE iterate
and the iteration count, exactly as solve.py:36-45 does), plus a ``maxiter``
guard: the reference loop has none and never terminates when the iteration
diverges.
"""
import numpy as np


def _gather(y_local, FF, group):
    """Local slab of the new iterate -> the full vector on every rank."""
    import torch
    import torch.distributed as dist
    if FF.shape[0] == FF.m_global:
        return y_local
    world = dist.get_world_size(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=y_local.device)
    sizes[dist.get_rank(group)] = y_local.numel()
    dist.all_reduce(sizes, group=group)
    sizes = sizes.tolist()
    width = max(sizes)
    pad = torch.zeros(width, dtype=y_local.dtype, device=y_local.device)
    pad[:y_local.numel()] = y_local
    out = torch.empty(world*width, dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r*width:r*width + sizes[r]] for r in range(world)])


def _allmax(v, FF, group):
    if FF.shape[0] == FF.m_global:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v if v == v else float('inf')], dtype=torch.float64, device=torch.device('cuda', FF.device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def solve_radiosity(FF, E, rho=1, albedo_placement='right', method='jacobi', tol=None,
                    maxiter=10000, group=None):
    import torch
    E = np.asarray(E)
    if tol is None:
        tol = np.finfo(E.dtype).resolution       # solve.py:6-8
        tol *= abs(E).max()
    if albedo_placement not in {'left', 'right'}:
        raise Exception('albedo_placement must be "left" or "right"')
    if method != 'jacobi':
        raise Exception('method must be jacobi on the device path (the reference\'s cg-right is '
                        '`assert False`, solve.py:108-109)')
    if FF.m_global != FF.shape[1]:
        raise Exception('radiosity needs a square form-factor matrix')
    dev = torch.device('cuda', FF.device)
    lo, hi = FF.row_start, FF.row_stop
    E_t = torch.as_tensor(np.ascontiguousarray(E, np.float64), device=dev)
    rho_arr = np.asarray(rho, np.float64)
    rho_t = None if rho_arr.ndim == 0 else torch.as_tensor(np.ascontiguousarray(rho_arr), device=dev)
    B = E_t.clone()
    niter = 0
    while True:
        niter += 1
        if albedo_placement == 'right':          # B1 = E + FF@(rho*B)            solve.py:41
            y, diff = FF.step(B, E_t[lo:hi].contiguous(), rho_t if rho_t is not None else float(rho_arr),
                              want_diff=True)
        else:                                    # B1 = E + rho*(FF@B)            solve.py:30
            y = FF.step(B)
            r = rho_t[lo:hi] if rho_t is not None else float(rho_arr)
            y = E_t[lo:hi] + r*y
            diff = float((y - B[lo:hi]).abs().max().item()) if y.numel() else 0.0
        diff = _allmax(diff, FF, group)
        if diff <= tol:
            break
        if not np.isfinite(diff) or niter >= maxiter:
            raise RuntimeError(f'Jacobi radiosity iteration did not converge after {niter} iterations '
                               f'(max change {diff}); the reference loop would not terminate (SURVEY P13)')
        B = _gather(y, FF, group)
    return B.cpu().numpy().astype(E.dtype, copy=False), niter


def compute_steady_state_temp(FF, E, rho, emiss, Fsurf=0., clamp=True, method='jacobi', group=None,
                              maxiter=10000):
    """Mirror of ``flux.model.compute_steady_state_temp`` for 1-D ``E``
    (reference src/flux/model.py:8-24) on a device-resident matrix."""
    import torch
    SIGMA_SB = 5.670374419e-8                    # scipy.constants.Stefan_Boltzmann
    E = np.asarray(E)
    if E.ndim != 1:
        raise Exception('E should be a vector on the device path')
    dev = torch.device('cuda', FF.device)
    B = solve_radiosity(FF, E, rho, 'right', method, maxiter=maxiter, group=group)[0]
    if clamp:
        B = np.maximum(0, B)
    v = torch.as_tensor(np.ascontiguousarray((1 - np.asarray(rho))*B + Fsurf, np.float64), device=dev)
    IR = _gather(FF.step(v), FF, group).cpu().numpy().astype(E.dtype, copy=False)    # FF@((1-rho)*B + Fsurf)
    Q = solve_radiosity(FF, IR, 1, 'right', method, maxiter=maxiter, group=group)[0]
    if clamp:
        Q = np.maximum(0, Q)
    tot = (1 - np.asarray(rho))*B + emiss*Q + Fsurf
    if clamp:
        tot = np.maximum(0, tot)
    return (tot/(emiss*SIGMA_SB))**0.25
