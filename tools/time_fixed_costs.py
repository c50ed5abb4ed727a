"""Where the rank-replicated (non-scaling) time of a sweep goes: a 500k-row shard (the 8-GPU share of
C3) on one GPU, wall-clock with synchronisation around each host-visible phase."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops
n, d, m = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
def sync(): torch.cuda.synchronize(); return time.perf_counter()
def sweep():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed', noise_var=0.1, chunk_rows=524288, deferred_checks=True).fit(X, y)
    _, C = est.estimator_.gradient_gram(want_G=False, check=False)
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n)
    est.estimator_.finish_checks()
    return tr.components_
for _ in range(3): sweep()
ts = []
for _ in range(10):
    t0 = sync(); sweep(); ts.append(sync() - t0)
print('sweep wall ms: min %.3f median %.3f' % (min(ts) * 1e3, sorted(ts)[5] * 1e3))
ops.start_timing()
for _ in range(10): sweep()
per = ops.stop_timing()
print({k: round(v[0] / 10, 3) for k, v in per.items()}, 'sum', round(sum(v[0] for v in per.values()) / 10, 3))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(10): sweep()
torch.cuda.synchronize(); pr.disable()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats('tottime').print_stats(40); print(sio.getvalue()[:9000])
