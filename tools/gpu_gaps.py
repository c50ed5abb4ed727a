"""Where one sweep on a 500k-row shard (the 8-GPU share of C3) spends its time: CUDA-event total of the sweep
through the public classes, the stage times the library brackets itself (edrgp_timing_begin / _end) and what is left
over -- the rank-replicated, latency-bound remainder R = total - (kuf + stats + gradients) that limits strong scaling."""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import dist as edist, model as emodel, ops

# under torchrun: every rank holds a shard of n rows (weak: n is PER RANK here) and the collectives are live
world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (500_000, 64, 512)))
g = torch.Generator(device='cuda').manual_seed(rank)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
B = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((d, 3)))[0], device='cuda')
y = torch.tanh(X @ B).sum(1) + 0.05 * torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z0 = X[:m].clone()
if world > 1:
    torch.distributed.broadcast(Z0, 0)
Z = Z0.cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))


def sweep():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True,
                                            method='fixed', noise_var=0.1, chunk_rows=524288, deferred_checks=True).fit(X, y)
    _, C = est.estimator_.gradient_gram(want_G=False, check=False, reduce=True)
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n * world)
    est.estimator_.finish_checks()
    return tr.components_


for _ in range(5):
    sweep()
torch.cuda.synchronize()
reps = 20
ops.start_timing()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(reps):
    sweep()
s1.record()
torch.cuda.synchronize()
st = ops.stop_timing()
total = s0.elapsed_time(s1) / reps
out = {'n': n, 'd': d, 'm': m, 'total_ms': total, 'stages_ms': {k: v[0] / reps for k, v in st.items()}}
scal = sum(out['stages_ms'].get(k, 0.0) for k in ('kuf', 'inducing_stats', 'grad_gram_cached'))
out['scalable_ms'] = scal
out['replicated_R_ms'] = total - scal
out['unattributed_ms'] = total - sum(out['stages_ms'].values())
# the same without the event pairs (they cost a little themselves)
s0.record()
for _ in range(reps):
    sweep()
s1.record()
torch.cuda.synchronize()
out['total_ms_untimed'] = s0.elapsed_time(s1) / reps
out['world'], out['rank'] = world, rank
if world > 1:
    # the collectives of a sweep on their own: CUDA events around back-to-back all-reduces of the three payloads
    coll = {}
    for name, count in (('table', 4 * world), ('stats', m * m + m + 1), ('gram', d * d)):
        t = torch.zeros(count, dtype=torch.float64, device='cuda')
        for _ in range(5):
            torch.distributed.all_reduce(t)
        torch.cuda.synchronize()
        s0.record()
        for _ in range(50):
            torch.distributed.all_reduce(t)
        s1.record()
        torch.cuda.synchronize()
        coll[name] = s0.elapsed_time(s1) / 50
    out['allreduce_ms_back_to_back'] = coll
for r in range(world):
    if r == rank:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
if world > 1:
    torch.distributed.destroy_process_group()
