"""GPU-side idle gaps between the timed ops of one sweep on a 500k-row shard (8-GPU share of C3)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops
n, d, m = 500_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
order = []
class T(ops._Timed):
    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True); self.e0.record()
    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); order.append((self.name, self.e0, e1)); return False
ops._Timed = T
def sweep():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed', noise_var=0.1, chunk_rows=524288, deferred_checks=True).fit(X, y)
    _, C = est.estimator_.gradient_gram(want_G=False, check=False)
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n)
    est.estimator_.finish_checks()
    return tr.components_
for _ in range(3): sweep()
torch.cuda.synchronize(); order.clear()
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
s0.record(); sweep(); s1.record(); torch.cuda.synchronize()
print('total %.3f ms' % s0.elapsed_time(s1))
prev = s0; pname = 'start'
for name, a, b in order:
    print('gap %-16s -> %-16s %.3f ms   | %-16s %.3f ms' % (pname, name, prev.elapsed_time(a), name, a.elapsed_time(b)))
    prev, pname = b, name
print('gap %-16s -> end %.3f ms' % (pname, prev.elapsed_time(s1)))
