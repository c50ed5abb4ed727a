"""The NVLink peer exchange of the composite sweep (csrc/peer.cu, edrgp_peer_* / edrgp_fixed_bind_peers /
edrgp_fixed_reduce_gram) on ONE GPU: several ranks are emulated inside this process -- every rank has its own workspace,
its own exchange buffer (plain device pointers stand in for the IPC mappings) and its own row shard -- and their composite
calls are issued phase by phase, so that every flag a kernel waits for has already been raised (a rank spinning for a
peer that shares its GPU would keep that peer from running).  What is checked is everything the exchange adds: the table
push, both copies of every payload over consecutive sweeps, the rank-order sums inside form_system_peer_kernel and
reduce_gram_peer_kernel, the flag / epoch bookkeeping per buffer set.  The real thing -- one process per GPU, buffers
mapped through cudaIpc handles -- is tools/check_multigpu_fixed.py under torchrun (profiles/r02_multigpu_fixed_peer_*.json)."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op          # noqa: E402  (the checker, never the product)

F64 = torch.float64
JITTER = 1e-8


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=F64, device='cuda')


class _Rank(object):
    def __init__(self, X, y, rank, world, m, chunk):
        from edrgp_b200 import ops
        self.X, self.y = _dev(X), _dev(y)
        n, d = X.shape
        self.fs = ops.FixedSweep(n, d, m, chunk, rank, world, 'cuda')
        self.K = torch.empty(n, m + (m & 1), dtype=F64, device='cuda')


def _sweep(ranks, Z, ell, sf2, beta, peer):
    """One sweep of all emulated ranks, phase by phase; returns per rank (P, byy, alpha, C)."""
    for r in ranks:
        r.fs.begin(r.X, r.y, Z, ell, sf2, r.K)
    torch.cuda.synchronize()
    if not peer and len(ranks) > 1:
        table = sum(r.fs.table for r in ranks)
        for r in ranks:
            r.fs.table.copy_(table)
    for r in ranks:
        r.fs.stats_pass(r.X, r.y, sf2, r.K, True)
    torch.cuda.synchronize()
    if not peer and len(ranks) > 1:
        stats = ranks[0].fs.stats.clone()
        for r in ranks[1:]:
            stats += r.fs.stats                      # rank order, as the peer kernels sum
        for r in ranks:
            r.fs.stats.copy_(stats)
    for r in ranks:
        r.fs.posterior(Z, sf2, JITTER, beta)
    torch.cuda.synchronize()
    for r in ranks:
        r.fs.grad(r.X, r.K, Z, ell, sf2, 1.0, r.fs.tail[3:4])
    torch.cuda.synchronize()
    out = []
    if peer:
        for r in ranks:
            r.fs.reduce_gram()
        torch.cuda.synchronize()
    elif len(ranks) > 1:
        C = ranks[0].fs.C.clone()
        for r in ranks[1:]:
            C += r.fs.C
        for r in ranks:
            r.fs.C.copy_(C)
    for r in ranks:
        out.append((r.fs.P.clone(), r.fs.byy.clone(), r.fs.alpha.clone(), r.fs.C.clone(), r.fs.tail.clone()))
    return out


@pytest.mark.parametrize("world,n,d,m", [(1, 3000, 10, 20), (2, 5001, 64, 128), (3, 4000, 32, 96)])
def test_emulated_ranks_exchange_equals_manual_reduction(world, n, d, m):
    from edrgp_b200 import _lib, ops
    lib = _lib.load()
    w = op.make_workload(n, d, m, seed=world)
    Z, ell = _dev(w['Z']), _dev(w['ell'])
    sf2, beta = 1.3, 10.0
    bounds = np.linspace(0, n, world + 1).astype(int)
    bounds[1:-1] += 1                                   # ragged shards
    chunk = 1024

    def make_ranks():
        return [_Rank(w['X'][bounds[r]:bounds[r + 1]], w['y'][bounds[r]:bounds[r + 1]], r, world, m, chunk)
                for r in range(world)]

    manual = make_ranks()
    ref = _sweep(manual, Z, ell, sf2, beta, peer=False)

    peers = make_ranks()
    off = (ctypes.c_int64 * 4)()
    nbytes = lib.edrgp_peer_layout(m, d, world, off)
    assert nbytes > 0
    bufs = []
    for _ in range(world):
        p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _lib.check(lib.edrgp_peer_alloc(nbytes, ctypes.byref(p), h), 'edrgp_peer_alloc')
        bufs.append(p.value)

    class _Ex(object):
        bases = (ctypes.c_void_p * world)(*bufs)
    try:
        for r in peers:
            r.fs.bind_peers(_Ex)
        # the Gram exchange publishes and consumes in ONE call (copy, signal, wait + sum), so a rank's call must not
        # run to completion before the later ranks have been issued: the ranks' calls go to separate streams and
        # overlap (a few small CTAs each), the phases before it are separated by device synchronisation
        streams = [torch.cuda.Stream() for _ in range(world)]
        for it in range(3):                              # both copies of every payload, epochs 1..3
            for r in peers:
                r.fs.begin(r.X, r.y, Z, ell, sf2, r.K)
            torch.cuda.synchronize()
            for r in peers:
                r.fs.stats_pass(r.X, r.y, sf2, r.K, True)
            torch.cuda.synchronize()
            for r in peers:
                r.fs.posterior(Z, sf2, JITTER, beta)
            torch.cuda.synchronize()
            for r in peers:
                r.fs.grad(r.X, r.K, Z, ell, sf2, 1.0, r.fs.tail[3:4])
            torch.cuda.synchronize()
            for r, s in zip(peers, streams):
                with torch.cuda.stream(s):                # small kernels (a few CTAs): the ranks' waits overlap
                    r.fs.reduce_gram()
            torch.cuda.synchronize()
            got = [(r.fs.P.clone(), r.fs.byy.clone(), r.fs.alpha.clone(), r.fs.C.clone(), r.fs.tail.clone()) for r in peers]
            for r in range(world):
                flag = ops.FixedSweep.decode_tail(got[r][4].cpu().numpy())[0]
                assert flag == 0, flag                   # no time-out bit, no non-finite rows
                for a, b in zip(got[r][:4], ref[r][:4]):
                    assert torch.equal(a, b)              # same sums in the same (rank) order: bit-identical
                for a, b in zip(got[r][:4], got[0][:4]):
                    assert torch.equal(a, b)              # and identical on every rank
    finally:
        for r in peers:
            r.fs.bind_peers(None)
        for b in bufs:
            lib.edrgp_peer_free(b)

    # against the oracle on all rows: the exchange adds nothing beyond summation order
    Pall = ref[0][0].cpu().numpy()
    ys = (w['y'] - w['y'].mean()) / w['y'].std()
    Kref = op.kuf_faithful(w['X'], w['Z'], w['ell'], sf2)
    assert np.max(np.abs(Pall - Kref.T.dot(Kref))) < 1e-11 * np.max(np.abs(Pall))
    assert np.max(np.abs(ref[0][1].cpu().numpy()[:m] - Kref.T.dot(ys))) < 1e-10 * np.max(np.abs(Kref.T.dot(ys)))


def test_reduce_gram_needs_a_bound_workspace():
    from edrgp_b200 import _lib, ops
    fs = ops.FixedSweep(1000, 8, 16, 512, 0, 1, 'cuda')
    with pytest.raises(_lib.EdrgpError):
        fs.reduce_gram()
