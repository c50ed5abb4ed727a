"""n-sharding plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the two
small reductions of the path.  SURVEY.md section 8(e).

The data rows are partitioned across ranks; Z, the hyper-parameters and alpha are replicated.
Per fixed-hyper-parameter sweep exactly two all-reduces run on the data path:
{P (m x m), b (m), y^T y, n} after the statistics pass and {C (d x d)} after the gradient pass
(plus 2d / 2 doubles for the scaler / target normaliser).  Everything else is rank-local.
"""
import contextlib
import os

import torch

_LOCAL_ONLY = False


@contextlib.contextmanager
def local_only():
    """Treat this process as a single-rank job inside the block (reference runs in multi-rank tests)."""
    global _LOCAL_ONLY
    prev, _LOCAL_ONLY = _LOCAL_ONLY, True
    try:
        yield
    finally:
        _LOCAL_ONLY = prev


def is_distributed():
    return (not _LOCAL_ONLY) and torch.distributed.is_available() and torch.distributed.is_initialized() \
        and torch.distributed.get_world_size() > 1


def world_size():
    return torch.distributed.get_world_size() if is_distributed() else 1


def rank():
    return torch.distributed.get_rank() if is_distributed() else 0


def allreduce_sum_(*tensors):
    """In-place sum over ranks of every tensor, packed into ONE collective."""
    if not is_distributed():
        return tensors
    if len(tensors) == 1 and tensors[0].is_contiguous():
        torch.distributed.all_reduce(tensors[0], op=torch.distributed.ReduceOp.SUM)   # in place, no staging copy
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
    off = 0
    for t in tensors:
        k = t.numel()
        t.copy_(flat[off:off + k].view_as(t))
        off += k
    return tensors


def allreduce_max_(tensor):
    """In-place maximum over ranks (failure flags)."""
    if is_distributed():
        torch.distributed.all_reduce(tensor, op=torch.distributed.ReduceOp.MAX)
    return tensor


def broadcast_(tensor, src=0):
    if is_distributed():
        torch.distributed.broadcast(tensor, src=src)
    return tensor


def shard_bounds(n, r=None, w=None):
    """Contiguous row block of rank r: X[lo:hi]."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def bind_to_local_cpus(device_index, min_cpus=2):
    """Pin this process to the CPUs NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root),
    so that the pinned host buffers it allocates AFTERWARDS (first touch) and the threads that feed the copy
    engine sit next to the GPU.  With eight ranks streaming their row shards from host memory the transfers
    otherwise cross the socket interconnect.  Does nothing -- and says so in the returned string -- when NVML or
    the affinity call is unavailable, when the local set is not a strict subset of the CPUs this process may
    already use, or when it has fewer than ``min_cpus`` CPUs.  Call it once per rank, before any pinned
    allocation; never raises."""
    if os.environ.get('EDRGP_BIND_LOCAL_CPUS', '1') == '0':
        return 'unchanged (EDRGP_BIND_LOCAL_CPUS=0)'
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError) as e:
        return 'unchanged (no sched_getaffinity: %s)' % e
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(device_index)
            if visible:                                   # NVML numbers the physical devices
                ids = [v.strip() for v in visible.split(',') if v.strip()]
                if idx < len(ids) and ids[idx].isdigit():
                    idx = int(ids[idx])
            handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = (max(allowed) // 64) + 1 if allowed else 1
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
    except Exception as e:                                # NVML missing / old driver / container without it
        return 'unchanged (NVML affinity unavailable: %s)' % type(e).__name__
    local = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
    target = local & allowed
    if len(target) < min_cpus or target == allowed:
        return 'unchanged (%d local of %d allowed CPUs)' % (len(target), len(allowed))
    try:
        os.sched_setaffinity(0, target)
    except OSError as e:
        return 'unchanged (sched_setaffinity: %s)' % e
    lo, hi = min(target), max(target)
    return 'bound to %d CPUs local to GPU %d (%d..%d)' % (len(target), int(device_index), lo, hi)
