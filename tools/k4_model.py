"""CPU model of the trace kernel's traversal: warp-level iteration counts per 32-ray batch for design
variants (see tools/k4_model.c).  No GPU needed.

    python tools/k4_model.py [grid_n] [rows]

Prints, per variant, the per-batch averages that set K4's instruction count (it is issue-bound):
phase A / B / C warp iterations, lane-level counts, deferred candidates, and an instruction estimate
    42*A + 50*B + 100*C + 130*flush + 620   (per-iteration costs from the SASS of the round-1 build,
                                              fixed part = cull share + compaction + ray set-up + target test)
to be compared with the measured 104 warp instructions per ray = 3330 per batch (profiles/r01b_kernels_ncu.md).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluxpy_b200 import meshes  # noqa: E402  (mesh generators only: pure NumPy)


def load():
    so = '/tmp/libk4model.so'
    src = os.path.join(ROOT, 'tools', 'k4_model.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', so, src, '-lm'])
    L = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    L.k4_build.restype = vp
    L.k4_build.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
    L.k4_count.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    L.k4_horizons.argtypes = [vp, ctypes.c_int, vp]
    L.k4_free.argtypes = [vp]
    L.k4_set_margin.argtypes = [ctypes.c_double]
    return L


def face_geometry(V, F):
    v0, v1, v2 = (V[F[:, k]].astype(np.float64) for k in range(3))
    P = (v0 + v1 + v2)/3
    C = np.cross(v1 - v0, v2 - v0)
    N = C/np.linalg.norm(C, axis=1)[:, None]
    return P, N


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 317
    nrows = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    body = len(sys.argv) > 3 and sys.argv[3] == 'body'
    if body:                                    # closed, cratered body (BASELINE config 4 stand-in), outward normals
        V, F = meshes.cratered_body(n, dtype=np.float32)
        P, N = face_geometry(V, F)
        N[(N*P).sum(1) < 0] *= -1
    else:
        V, F = meshes.gaussian_crater(n, 0, dtype=np.float32)
        P, N = face_geometry(V, F)
        N[N[:, 2] < 0] *= -1                    # upward normals, as the bench's meshes.upward_normals
    V = np.ascontiguousarray(V, np.float32)
    F32 = np.ascontiguousarray(F, np.int32)
    P, N = np.ascontiguousarray(P), np.ascontiguousarray(N)
    L = load()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    M = L.k4_build(len(V), p(V), len(F32), p(F32), p(P), p(N))
    nf = len(F32)
    rows = np.ascontiguousarray(np.linspace(0, nf - 1, nrows).astype(np.int32))
    print(f'{"cratered body" if body else "G"}({n}): {nf} faces, {nrows} sample rows x all columns')
    hdr = ('variant', 'rays/b', 'list', 'A', 'B', 'C', 'A hits', 'B lane', 'C lane', 'cand', 'flush', 'est instr/batch',
           'instr/ray')
    print(' | '.join(hdr))
    L.k4_set_margin(float(os.environ.get('K4_MARGIN', '2e-3')))
    print('horizon margin (sine):', os.environ.get('K4_MARGIN', '2e-3'))
    hors = {}
    for zl in (64, 256, 1024, 4096):
        hors[zl] = np.zeros(nf, np.float32)
        L.k4_horizons(M, zl, p(hors[zl]))
        h = hors[zl][np.isfinite(hors[zl])]
        print(f'horizon of the {zl}-leaf zone: median sin(elev) {np.median(h):.3f}, 90 % {np.quantile(h, 0.9):.3f}, '
              f'unbounded {100*(1 - len(h)/nf):.1f} %')
    for name, chunk, bf, axes, expand, zl in (
            ('chunk 1024 (as built)', 1024, 0, 0, 0, 0), ('chunk 1024 + per-batch filter', 1024, 1, 0, 0, 0),
            ('chunk 1024 + 2 shaft axes', 1024, 0, 1, 0, 0), ('chunk 512', 512, 0, 0, 0, 0),
            ('expand records > 4096 leaves', 1024, 0, 0, 4096, 0),
            ('horizon skip, 64-leaf zones', 1024, 0, 0, 0, 64), ('horizon skip, 256-leaf zones', 1024, 0, 0, 0, 256),
            ('horizon skip, 1024-leaf zones', 1024, 0, 0, 0, 1024), ('horizon skip, 4096-leaf zones', 1024, 0, 0, 0, 4096),
            ('horizon 256 + expand > 4096', 1024, 0, 0, 4096, 256)):
        out = np.zeros(32)
        L.k4_count(M, len(rows), p(rows), chunk, 1e-5, bf, axes, expand, zl, p(hors[zl]) if zl else None, p(out))
        b = out[0]
        A, B, C, fl = out[2]/b, out[3]/b, out[4]/b, out[9]/b
        if bf:
            A = out[13]/b
        # the per-batch filter costs ~120 instructions per batch; a unit's list build + cull are shared
        # by its batches (chunk / 1024 scales how many batches share them)
        est = 42*A + 50*B + 100*C + 130*fl + 620 + (120 if bf else 0)
        print(f'{name} | {out[1]/b:.1f} | {out[11]/out[10]:.1f}->{out[12]/out[10]:.1f} | {A:.2f} | {B:.2f} | {C:.2f} | '
              f'{out[5]/b:.1f} | {out[6]/out[1]:.2f} | {out[7]/out[1]:.2f} | {out[8]/out[1]:.2f} | {fl:.2f} | {est:.0f} | '
              f'{est/(out[1]/b):.0f}   (units with common ancestor {100*out[15]/out[10]:.0f} %, survivors {100*out[1]/out[14]:.0f} %)')
        if zl:
            print(f'    batches leaving above the source horizon {100*out[24]/b:.0f} %, rays arriving above the target horizon {100*out[23]/out[1]:.0f} %')
        if name.endswith('(as built)'):
            r = out[1]
            print(f'    per ray: phase-A hits {out[5]/r:.2f} ({out[21]/r:.2f} on records of <= 64 leaves), phase-B hits {out[22]/r:.2f}; '
                  f'phase-C visits rooted in A {out[16]/r:.2f}, in B {out[17]/r:.2f}; deferred candidates {out[8]/r:.2f}: '
                  f'the source triangle itself {out[18]/r:.2f}, within 64 leaves of the source {out[19]/r:.2f}, of the target {out[20]/r:.2f}')
    L.k4_free(M)


if __name__ == '__main__':
    main()
