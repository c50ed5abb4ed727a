"""n-sharding plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the two
small reductions of the path.  SURVEY.md section 8(e).

The data rows are partitioned across ranks; Z, the hyper-parameters and alpha are replicated.
Per fixed-hyper-parameter sweep exactly two all-reduces run on the data path:
{P (m x m), b (m), y^T y, n} after the statistics pass and {C (d x d)} after the gradient pass
(plus 2d / 2 doubles for the scaler / target normaliser).  Everything else is rank-local.
"""
import contextlib

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
    flat = torch.cat([t.reshape(-1) for t in tensors])
    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
    off = 0
    for t in tensors:
        k = t.numel()
        t.copy_(flat[off:off + k].view_as(t))
        off += k
    return tensors


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
