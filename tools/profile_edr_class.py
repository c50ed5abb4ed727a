"""cProfile of EffectiveDimensionalityReduction.fit from pinned host rows at the headline shape (development aid)."""
import cProfile, pstats, sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel
n, d, m = 4_000_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
Xh = torch.empty(n, d, dtype=torch.float64, pin_memory=True); Xh.copy_(X)
yh = torch.empty(n, dtype=torch.float64, pin_memory=True); yh.copy_(y)
Xn, yn = Xh.numpy(), yh.numpy()
del X, y
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
def fit():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed',
                                            noise_var=0.1, chunk_rows=524288, deferred_checks=True)
    return eb.EffectiveDimensionalityReduction(est, eb.GramEighTransformer(), n_components=None, normalize=True,
                                               keep_gradients=False).fit(Xn, yn)
for _ in range(2): fit()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); fit(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({'edr_fit_ms': ts}))
pr = cProfile.Profile(); pr.enable(); fit(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
