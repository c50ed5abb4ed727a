"""Host -> device bandwidth of this box, one rank per GPU, all ranks at the same time (run under torchrun, or alone):

    pinned    cudaMemcpyAsync of one pinned buffer per rank (the ceiling of the end-to-end arm of bench.py)
    staged    the library's streamer (edrgp_h2d_*) from an ordinary NumPy array: host threads -> pinned ring -> DMA
    pageable  torch's copy from the same NumPy array (what `tensor.cuda()` does)

Prints one JSON line on rank 0: per-rank GB/s (min / mean / max over ranks) and the aggregate.
"""
import ctypes
import json
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edrgp_b200 import _lib, gp_model                       # noqa: E402

world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nbytes = mb << 20
rows, rb = nbytes // 512, 512
lib = _lib.load()
dst = torch.empty(nbytes // 8, dtype=torch.float64, device='cuda')
pinned = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
pinned.fill_(1.0)
page = np.ones(nbytes // 8)
threads = gp_model._host_copy_threads()


def barrier():
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    best = 0.0
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best


def streamer(src_ptr):
    st = torch.cuda.current_stream().cuda_stream
    h = lib.edrgp_h2d_open(src_ptr, dst.data_ptr(), rows, rb, rb, 131072, threads, 3, st, 0, 0, 0)
    assert h, lib.edrgp_last_error()
    _lib.check(lib.edrgp_h2d_wait(h, rows, rows, st), 'wait')
    lib.edrgp_h2d_close(h)


res = {'pinned_memcpy': timed(lambda: dst.copy_(pinned, non_blocking=True)),
       'pinned_streamer': timed(lambda: streamer(pinned.data_ptr())),
       'staged_streamer': timed(lambda: streamer(page.ctypes.data)),
       'pageable_torch': timed(lambda: dst.copy_(torch.from_numpy(page)))}
t0 = time.perf_counter()
for _ in range(3):
    np.copyto(pinned.numpy(), page)
res['host_memcpy_1thread'] = 3 * nbytes / (time.perf_counter() - t0) / 1e9
vals = torch.tensor([res[k] for k in sorted(res)], dtype=torch.float64, device='cuda')
if world > 1:
    allv = [torch.empty_like(vals) for _ in range(world)]
    torch.distributed.all_gather(allv, vals)
    allv = torch.stack(allv).cpu().numpy()
else:
    allv = vals.cpu().numpy()[None]
if rank == 0:
    out = {'world': world, 'megabytes_per_rank': mb, 'copy_threads_per_rank': threads, 'cpus': len(os.sched_getaffinity(0))}
    for i, k in enumerate(sorted(res)):
        out[k] = {'per_rank_min': float(allv[:, i].min()), 'per_rank_mean': float(allv[:, i].mean()),
                  'per_rank_max': float(allv[:, i].max()), 'aggregate_GBps': float(allv[:, i].sum())}
    print(json.dumps(out))
if world > 1:
    torch.distributed.destroy_process_group()
