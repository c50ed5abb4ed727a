"""Device-resident fixed-hyper-parameter sweep time for the BASELINE.json config shapes
(configs other than C3 are parity-test cases; this records what they cost on one B200)."""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops

CONFIGS = {'C1': (500, 10, 20), 'C2': (100_000, 32, 256), 'C3': (4_000_000, 64, 512),
           'C4': (1_000_000, 512, 1024), 'C5_per_gpu': (2_000_000, 128, 2048)}
args = [a for a in sys.argv[1:] if not a.startswith('--')]
precisions = ['fp64', 'tf32x3'] if '--both' in sys.argv else ['fp64']
which = args or list(CONFIGS)
for name, precision in [(nm, pr) for nm in which for pr in precisions]:
    n, d, m = CONFIGS[name]
    if precision == 'tf32x3' and d > 64:
        continue
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    B = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((d, 3)))[0], device='cuda')
    y = torch.tanh(X @ B).sum(1) + 0.05 * torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
    Z = X[:m].cpu().numpy()
    ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))

    def sweep():
        est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True,
                                                method='fixed', noise_var=0.1, chunk_rows=262144, precision=precision).fit(X, y)
        _, C = est.estimator_.gradient_gram(want_G=False)
        return eb.GramEighTransformer(n_components=3).fit_gram(C, n).components_
    for _ in range(2):
        comps = sweep()
    torch.cuda.synchronize()
    ts = []
    ops.start_timing()
    for _ in range(3):
        t0 = time.perf_counter(); sweep(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    per = ops.stop_timing()
    ms = min(ts) * 1e3
    flop = n * (6.0 * m * d + 2.0 * m * m + 2.0 * d * d + 3.0 * m)
    print(json.dumps({'config': name, 'precision': precision, 'n': n, 'd': d, 'm': m, 'ms': ms, 'pts_per_s': n / ms * 1e3,
                      'alg_tflops': flop / ms / 1e9, 'stages_ms': {k: v[0] / 3 for k, v in per.items()}}))
    del X, y
    torch.cuda.empty_cache()
