"""Where the end-to-end sweep (pinned host rows in) spends its time: per-op CUDA-event times and the idle gaps between them."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops
n, d, m = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
Xh = torch.empty(n, d, dtype=torch.float64, pin_memory=True); yh = torch.empty(n, dtype=torch.float64, pin_memory=True)
Xh.copy_(X); yh.copy_(y); torch.cuda.synchronize()
Xnp, ynp = Xh.numpy(), yh.numpy()
del X, y
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
order = []
class T(ops._Timed):
    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True); self.e0.record()
    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); order.append((self.name, self.e0, e1)); return False
ops._Timed = T
def sweep():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed', noise_var=0.1, chunk_rows=524288, deferred_checks=True).fit(Xnp, ynp)
    _, C = est.estimator_.gradient_gram(want_G=False, check=False)
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n)
    est.estimator_.finish_checks()
    return tr.components_
for _ in range(3): sweep()
torch.cuda.synchronize(); order.clear()
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s0.record(); sweep(); s1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print('total %.3f ms (wall %.3f)' % (s0.elapsed_time(s1), (t1 - t0) * 1e3))
prev = s0; pname = 'start'; busy = 0.0; gaps = 0.0
for name, a, b in order:
    gap = prev.elapsed_time(a); dur = a.elapsed_time(b); busy += dur; gaps += gap
    if gap > 0.15 or name not in ('kuf', 'inducing_stats'):
        print('gap %-16s -> %-16s %.3f ms   | %-16s %.3f ms' % (pname, name, gap, name, dur))
    prev, pname = b, name
print('gap %-16s -> end %.3f ms' % (pname, prev.elapsed_time(s1)))
print('busy %.3f gaps %.3f' % (busy, gaps))
