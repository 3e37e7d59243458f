import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import fluxpy_b200
from fluxpy_b200 import meshes, form_factors, _lib
# raw PCIe
x = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f'D2H pinned 1 GiB: {1.0737/dt:.1f} GB/s')
hp = torch.empty(1 << 30, dtype=torch.uint8)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); hp.copy_(x); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f'D2H pageable 1 GiB: {1.0737/dt:.1f} GB/s')
V, F = meshes.gaussian_crater(317, 0, dtype=np.float32)
sm = fluxpy_b200.CudaTrimeshShapeModel(V, F, meshes.upward_normals(V, F))
nf = sm.num_faces
for sub, bps in ((512, 3), (512, 4), (256, 3), (1024, 3), (1024, 4)):
    sm.set_option('sub_rows', sub); sm.set_option('blocks_per_sm', bps)
    for rep in range(4):
        I = np.arange(4096) + 4096*(rep + 3)
        t = time.perf_counter(); sm._sync_face_data(); t1 = time.perf_counter()
        FF = fluxpy_b200.get_form_factor_matrix(sm, I); t2 = time.perf_counter()
        st = dict(form_factors.last_stats)
        print(f'bps={bps} sub={sub} rep={rep} sync_face={1e3*(t1-t):.1f} ms  call={1e3*(t2-t1):.1f} ms  trace={st["ms_trace"]:.1f} copy_stream_span={st["ms_fill"]:.1f} prep={st["ms_prepare"]:.1f} nnz={st["nnz"]:.3e} free_blocks={len(_lib.arena.free)} retries={type(sm).overflow_retries}', flush=True)
        del FF
    t = time.perf_counter(); m, n, c, st = sm._ff_assemble_device(I, None, 1e-5); t2 = time.perf_counter()
    print(f'   device-resident call={1e3*(t2-t):.1f} ms trace={st.ms_trace:.1f} span={st.ms_fill:.1f}')
