"""Concurrent pinned D2H bandwidth on every GPU of the box (one process per GPU via torchrun)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get('LOCAL_RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
bind = os.environ.get('BIND', '0') == '1'
if bind:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(rank)
    try:
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception as e:
        print('bind failed', e)
x = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for rep in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): h.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f'rank {rank} bind={bind} cpus={sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))}) D2H {4*1.0737/dt:.1f} GB/s', flush=True)
if world > 1: dist.destroy_process_group()
