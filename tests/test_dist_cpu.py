"""Multi-rank host logic on CPU: two gloo processes (127.0.0.1).  Checks the packed all-reduce and
broadcast helpers, the contiguous row sharding, and the shard-sum algebra of the path (SURVEY.md
section 4(iv), 8(e)): statistics {P, b, yy} and the Gram matrix C reduced over row shards equal the
single-process result, and every rank derives the same alpha and EDR directions from them.  The
per-shard arithmetic here is the oracle's (no GPU in this container); the CUDA path performs the
same two reductions through the same helpers (edrgp_b200/dist.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp      # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as tdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    tdist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from edrgp_b200 import dist
        from oracle import pipeline as op
        assert dist.is_distributed() and dist.world_size() == world and dist.rank() == rank
        # packed all-reduce of tensors of different shapes
        a = torch.full((3, 2), float(rank + 1), dtype=torch.float64)
        b = torch.arange(5, dtype=torch.float64) * (rank + 1)
        dist.allreduce_sum_(a, b)
        assert torch.equal(a, torch.full((3, 2), 3.0, dtype=torch.float64))
        assert torch.equal(b, torch.arange(5, dtype=torch.float64) * 3)
        z = torch.full((4,), float(rank), dtype=torch.float64)
        dist.broadcast_(z, 0)
        assert torch.equal(z, torch.zeros(4, dtype=torch.float64))
        # a single contiguous tensor is reduced in place (no staging copy); a failure flag by maximum
        one = torch.full((6,), float(rank + 1), dtype=torch.float64)
        view = one[1:5]
        dist.allreduce_sum_(view)
        assert one.tolist() == [rank + 1.0, 3.0, 3.0, 3.0, 3.0, rank + 1.0]
        flag = torch.tensor([float(rank == 1)], dtype=torch.float64)
        dist.allreduce_max_(flag)
        assert float(flag[0]) == 1.0
        # the NVLink peer exchange is a CUDA / NCCL route: a gloo job gets None on every rank and keeps the all-reduces
        assert dist.peer_exchange(512, 64) is None
        # the targets' moments table of the composite sweep (csrc/sweep.cu): every rank fills its own row
        # [n_r, pivot_r, sum (y - pivot_r), sum (y - pivot_r)^2]; one summed table gives the GLOBAL mean and std
        rng = np.random.RandomState(7)
        y_all = 1e3 + 2.5 * rng.standard_normal(5001)
        lo_y, hi_y = dist.shard_bounds(y_all.size)
        ys_ = y_all[lo_y:hi_y]
        table = torch.zeros(world, 4, dtype=torch.float64)
        piv = ys_[0]
        table[rank] = torch.tensor([ys_.size, piv, np.sum(ys_ - piv), np.sum((ys_ - piv) ** 2)])
        dist.allreduce_sum_(table)
        t = table.numpy()
        N = t[:, 0].sum()
        mean = np.sum(t[:, 0] * t[:, 1] + t[:, 2]) / N
        m2 = np.sum(t[:, 3] + 2 * (t[:, 1] - mean) * t[:, 2] + t[:, 0] * (t[:, 1] - mean) ** 2)
        assert abs(mean - y_all.mean()) < 1e-13 * abs(y_all.mean())
        assert abs(np.sqrt(m2 / N) - y_all.std()) < 1e-12 * y_all.std()

        n, d, m = 3001, 6, 25
        w = op.make_workload(n, d, m, seed=2, k_true=2)
        lo, hi = dist.shard_bounds(n)
        Xs, ys = w['X'][lo:hi], w['y'][lo:hi]
        P, bb, yy = op.inducing_stats_chunked(Xs, ys, w['Z'], w['ell'], w['sf2'], chunk=512)
        Pt, bt, yt = torch.from_numpy(P), torch.from_numpy(bb), torch.tensor([yy], dtype=torch.float64)
        dist.allreduce_sum_(Pt, bt, yt)
        sol = op.solve_from_stats(op.kuu(w['Z'], w['ell'], w['sf2']), Pt.numpy(), bt.numpy(), float(yt[0]), n,
                                  w['sf2'], w['noise'])
        C = torch.from_numpy(op.grad_gram_chunked(Xs, w['Z'], w['ell'], w['sf2'], sol['alpha'], chunk=512))
        dist.allreduce_sum_(C)
        comps, lam, _ = op.edr_from_gram(C.numpy(), 2)
        out[rank] = {'lo': lo, 'hi': hi, 'alpha': sol['alpha'], 'bound': sol['bound'], 'C': C.numpy(), 'comps': comps}
    finally:
        tdist.destroy_process_group()


def test_two_rank_shard_sum_equals_single_process():
    from oracle import pipeline as op
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    n, d, m = 3001, 6, 25
    assert (r0['lo'], r0['hi'], r1['lo'], r1['hi']) == (0, 1501, 1501, 3001)
    # every rank ends with the same reduced quantities
    assert np.array_equal(r0['C'], r1['C']) and np.array_equal(r0['alpha'], r1['alpha'])
    w = op.make_workload(n, d, m, seed=2, k_true=2)
    P, b, yy = op.inducing_stats_chunked(w['X'], w['y'], w['Z'], w['ell'], w['sf2'])
    sol = op.solve_from_stats(op.kuu(w['Z'], w['ell'], w['sf2']), P, b, yy, n, w['sf2'], w['noise'])
    C = op.grad_gram_chunked(w['X'], w['Z'], w['ell'], w['sf2'], sol['alpha'])
    assert abs(r0['bound'] - sol['bound']) < 1e-10 * abs(sol['bound'])
    assert np.max(np.abs(r0['C'] - C)) < 1e-9 * np.max(np.abs(C))
    comps, _, _ = op.edr_from_gram(C, 2)
    assert op.principal_angle(r0['comps'], comps) < 1e-6


def test_shard_bounds_cover_rows_exactly():
    from edrgp_b200 import dist
    for n in (0, 1, 7, 4_000_000, 4_000_001):
        for w in (1, 2, 4, 8):
            b = [dist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _fake_nvml(mask_words):
    import types
    m = types.ModuleType('pynvml')
    m.nvmlInit = lambda: None
    m.nvmlShutdown = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: ('gpu', i)
    m.nvmlDeviceGetCpuAffinity = lambda h, n: list(mask_words)[:n] + [0] * max(0, n - len(mask_words))
    return m


def test_bind_to_local_cpus_guards(monkeypatch):
    """Rank placement next to the GPU (multi-rank bench runs): binds only to a usable strict subset of the CPUs
    the process may already use, reports what it did, never raises."""
    import os
    import sys
    from edrgp_b200 import dist
    calls = []
    monkeypatch.setattr(os, 'sched_getaffinity', lambda pid: set(range(8)), raising=False)
    monkeypatch.setattr(os, 'sched_setaffinity', lambda pid, cpus: calls.append(set(cpus)), raising=False)
    monkeypatch.delenv('CUDA_VISIBLE_DEVICES', raising=False)
    monkeypatch.delenv('EDRGP_BIND_LOCAL_CPUS', raising=False)
    # CPUs 4..7 are local: bound
    monkeypatch.setitem(sys.modules, 'pynvml', _fake_nvml([0xF0]))
    msg = dist.bind_to_local_cpus(1)
    assert calls == [{4, 5, 6, 7}] and msg.startswith('bound to 4 CPUs local to GPU 1')
    # every allowed CPU is local: nothing to do
    calls.clear()
    monkeypatch.setitem(sys.modules, 'pynvml', _fake_nvml([0xFF]))
    assert dist.bind_to_local_cpus(0).startswith('unchanged') and not calls
    # a single local CPU would serialise the rank's threads: refused
    monkeypatch.setitem(sys.modules, 'pynvml', _fake_nvml([0x01]))
    assert dist.bind_to_local_cpus(0).startswith('unchanged') and not calls
    # local CPUs outside the allowed set (cgroup / taskset): refused
    monkeypatch.setitem(sys.modules, 'pynvml', _fake_nvml([0xFF00]))
    assert dist.bind_to_local_cpus(0).startswith('unchanged') and not calls
    # NVML failing in any way: unchanged, no exception
    bad = _fake_nvml([0xF0])
    bad.nvmlInit = lambda: (_ for _ in ()).throw(RuntimeError('no driver'))
    monkeypatch.setitem(sys.modules, 'pynvml', bad)
    assert dist.bind_to_local_cpus(0).startswith('unchanged') and not calls
    # switched off
    monkeypatch.setenv('EDRGP_BIND_LOCAL_CPUS', '0')
    monkeypatch.setitem(sys.modules, 'pynvml', _fake_nvml([0xF0]))
    assert dist.bind_to_local_cpus(0).startswith('unchanged') and not calls
