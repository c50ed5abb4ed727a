"""n-sharding plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the two
small reductions of the path.  SURVEY.md section 8(e).

The data rows are partitioned across ranks; Z, the hyper-parameters and alpha are replicated.
Per fixed-hyper-parameter sweep three small reductions run on the data path: the targets' moments table,
{P (m x m), b (m), y^T y} after the statistics pass and {C (d x d)} after the gradient pass.  On an NVLink box
they run INSIDE the library's kernels over peer memory (``peer_exchange``, csrc/peer.cu); otherwise, and for
everything off the composite sweep (scaler, optimisation path), as ``torch.distributed`` all-reduces.
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


# --- NVLink peer exchange (csrc/peer.cu): the sweep's three small reductions without a library collective -----------
_PEER = {}            # (m, d, world) -> PeerExchange, or False when the ranks could not map each other's buffers


class PeerExchange(object):
    """This rank's exchange buffer (cudaMalloc + IPC handle) and the peers' buffers mapped into this process.
    Built COLLECTIVELY (every rank calls ``peer_exchange`` with the same shape at the same point of the program);
    ``bases`` is what ``edrgp_fixed_bind_peers`` takes."""

    def __init__(self, m, d):
        import ctypes
        from . import _lib
        lib = _lib.load()
        self.m, self.d, self.world, self.rank = int(m), int(d), world_size(), rank()
        self._lib, self._own, self._opened = lib, None, []
        off = (ctypes.c_int64 * 4)()
        nbytes = lib.edrgp_peer_layout(self.m, self.d, self.world, off)
        ok = nbytes > 0
        handle = ctypes.create_string_buffer(64)
        own = ctypes.c_void_p()
        if ok:
            ok = lib.edrgp_peer_alloc(nbytes, ctypes.byref(own), handle) == 0
        if ok:
            self._own = own.value
        handles = [None] * self.world
        torch.distributed.all_gather_object(handles, handle.raw if ok else None)
        ptrs = [None] * self.world
        ok = ok and all(h is not None for h in handles)
        if ok:
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs[r] = self._own
                    continue
                p = ctypes.c_void_p()
                if lib.edrgp_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                    ok = False
                    break
                ptrs[r] = p.value
                self._opened.append(p.value)
        # every rank or none: a rank that could not map a peer takes everybody back to the library collectives
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        self.ok = bool(flag.item() > 0.5)
        self.bases = (ctypes.c_void_p * self.world)(*ptrs) if self.ok else None
        if not self.ok:
            self.close()

    def close(self):
        for p in self._opened:
            self._lib.edrgp_peer_close(p)
        self._opened = []
        if self._own is not None:
            self._lib.edrgp_peer_free(self._own)
            self._own = None


def peer_exchange(m, d):
    """The exchange buffers of this job for an (m, d) sweep, or None: single process, a CPU / gloo job, more than 16
    ranks, ``EDRGP_COLLECTIVES=nccl`` in the environment, or a rank that could not map a peer's buffer (no NVLink /
    PCIe peer access).  COLLECTIVE on first use per shape; every rank gets the same answer."""
    if not is_distributed() or world_size() > 16 or os.environ.get('EDRGP_COLLECTIVES', 'peer') == 'nccl':
        return None
    if torch.distributed.get_backend() != 'nccl' or not torch.cuda.is_available():
        return None
    key = (int(m), int(d), world_size())
    ex = _PEER.get(key)
    if ex is None:
        ex = PeerExchange(m, d)
        if not ex.ok:
            import warnings
            warnings.warn("edrgp_b200: the ranks could not map each other's exchange buffers; the sweep's reductions "
                          "run as torch.distributed all-reduces", RuntimeWarning)
            ex = False
        _PEER[key] = ex
    return ex or None


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
