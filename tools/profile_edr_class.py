import sys, time, json, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import ops
n, d, m = 4_000_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g) + 0.5
y = torch.tanh(X[:, 0]) + 0.05 * torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
def fit():
    np.random.seed(0)
    return eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor(kernels='RBF', kernel_options={'ARD': True, 'lengthscale': 8.0},
                                          num_inducing=m, method='fixed', noise_var=0.1, chunk_rows=524288),
        eb.GramEighTransformer(), n_components=None, normalize=True, keep_gradients=False).fit(X, y)
for _ in range(2): fit()
torch.cuda.synchronize()
ops.start_timing()
t0 = time.perf_counter(); fit(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
per = ops.stop_timing()
print('wall ms', wall * 1e3, {k: round(v[0], 2) for k, v in per.items()}, 'sum', round(sum(v[0] for v in per.values()), 1))
pr = cProfile.Profile(); pr.enable(); fit(); torch.cuda.synchronize(); pr.disable()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats('tottime').print_stats(14); print(sio.getvalue()[:3000])
