"""Brute-force check of the "horizon skip" (profiles/r01b_k4_model.md): exact per-face horizons, then every
sampled ray that clears a horizon is Pluecker-tested (float32, as the kernel does) against every triangle
of the zone it would skip.  Any hit is a violation.  CPU only.

    python tools/k4_horizon_check.py [grid_n] [rows] [zone_leaves] [c_pert] [scale] [formula]
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluxpy_b200 import meshes  # noqa: E402
from tools.k4_model import face_geometry  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 159
    nrows = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    zone = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    c_pert = float(sys.argv[4]) if len(sys.argv) > 4 else 16.0
    scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    formula = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    so = '/tmp/libk4model_chk.so'
    src = os.path.join(ROOT, 'tools', 'k4_model.c')
    subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', so, src, '-lm'])
    L = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    L.k4_build.restype = vp
    L.k4_build.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
    L.k4_horizons_exact.argtypes = [vp, ctypes.c_int, ctypes.c_double, vp]
    L.k4_check_horizon.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_double, vp]
    V, F = meshes.gaussian_crater(n, 0, dtype=np.float32)
    V = np.ascontiguousarray(V*np.float32(scale), np.float32)
    P, N = face_geometry(V, F)
    N[N[:, 2] < 0] *= -1
    # the shape model's own P, N are float32 arrays: use their float32 values, as the device does
    P = np.ascontiguousarray(P.astype(np.float32).astype(np.float64))
    N = np.ascontiguousarray(N.astype(np.float32).astype(np.float64))
    F32 = np.ascontiguousarray(F, np.int32)
    p = lambda a: a.ctypes.data_as(vp)
    M = L.k4_build(len(V), p(V), len(F32), p(F32), p(P), p(N))
    nf = len(F32)
    hor = np.zeros(nf, np.float32)
    L.k4_set_formula(formula)
    L.k4_horizons_exact(M, zone, c_pert, p(hor))
    fin = hor[np.isfinite(hor)]
    print(f'G({n},0) x {scale}: {nf} faces, zone {zone} leaves, perturbation {c_pert} ulp: horizon median {np.median(fin):.4f}, '
          f'90 % {np.quantile(fin, 0.9):.4f}, unbounded {100*(1 - len(fin)/nf):.2f} %')
    rows = np.ascontiguousarray(np.linspace(0, nf - 1, nrows).astype(np.int32))
    out = np.zeros(8)
    L.k4_check_horizon(M, len(rows), p(rows), zone, p(hor), 1e-5, p(out))
    print(f'rays {out[0]:.0f} (own target missed {out[6]:.0f}); clear the source horizon {100*out[1]/max(out[0] - out[6], 1):.1f} %, '
          f'the target horizon {100*out[2]/max(out[0] - out[6], 1):.1f} %; Pluecker tests {out[5]:.3g}; '
          f'VIOLATIONS source end {out[3]:.0f}, target end {out[4]:.0f}')


if __name__ == '__main__':
    main()
