"""Cost of one VFE objective + gradient evaluation (what L-BFGS pays per iteration) on one GPU."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import model as emodel, ops
args = [a for a in sys.argv[1:] if not a.startswith('--')]
precision = 'tf32x3' if '--tf32' in sys.argv else 'fp64'
n, d, m = (int(a) for a in (args[:3] if len(args) > 2 else (4_000_000, 64, 512)))
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
B = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((d, 3)))[0], device='cuda')
y = torch.tanh(X @ B).sum(1) + 0.05 * torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
mod = emodel.SparseGPRegression(X, y[:, None], kernel=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True,
                                chunk_rows=524288, noise_var=0.1, precision=precision)
x0 = mod._get_optimizer_array()
mod._need_grad = True
for _ in range(2):
    mod._objective_grads(x0)
torch.cuda.synchronize()
ops.start_timing()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); f, gr = mod._objective_grads(x0); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
per = ops.stop_timing()
print(json.dumps({'precision': precision, 'n': n, 'd': d, 'm': m, 'eval_ms': min(ts) * 1e3, 'objective': f, 'grad_norm': float(np.linalg.norm(gr)),
                  'timed_ops_ms': {k: v[0] / 3 for k, v in per.items()}}))
t0 = time.perf_counter(); mod.optimize(max_iters=10); torch.cuda.synchronize()
print(json.dumps({'optimize_10_iters_s': time.perf_counter() - t0, 'evals': mod.optimization_runs[-1][2]['funcalls'],
                  'f_end': mod.optimization_runs[-1][0]}))
