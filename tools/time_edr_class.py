"""The public orchestrator at the headline shape: EffectiveDimensionalityReduction.fit from host rows
(scaler + sweep + projection + last fit on the projected rows), fixed hyper-parameters."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel
n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (4_000_000, 64, 512)))
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g) * (1 + torch.arange(d, device='cuda') / d) + 0.5
B = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((d, 3)))[0], device='cuda')
y = torch.tanh((X - 0.5) @ B).sum(1) + 0.05 * torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Xh = torch.empty(n, d, dtype=torch.float64, pin_memory=True); Xh.copy_(X)
yh = torch.empty(n, dtype=torch.float64, pin_memory=True); yh.copy_(y)
Xn, yn = Xh.numpy(), yh.numpy()
Xs0 = ((X[:m] - X.mean(0)) / X.std(0, unbiased=False)).cpu().numpy()      # inducing inputs in the scaled space
Z0 = Xs0
ell = np.sqrt(d) * np.ones(d)
USE_Z = '--draw-z' not in sys.argv     # Z=None follows GPy: np.random.permutation(n)[:m], ~55 ms of host time at n = 4M
def fit(Xa, ya, k):
    np.random.seed(0)
    edr = eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor(kernels='RBF', kernel_options={'ARD': True, 'lengthscale': float(np.sqrt(d))},
                                          num_inducing=m, method='fixed', noise_var=0.1, chunk_rows=524288,
                                          Z=(None if not USE_Z else Z0)),
        eb.GramEighTransformer(), n_components=k, normalize=True, keep_gradients=False)
    return edr.fit(Xa, ya)
for name, Xa, ya in (('device rows', X, y), ('host rows', Xn, yn)):
    for _ in range(2): fit(Xa, ya, None)
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); edr = fit(Xa, ya, None); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    from edrgp_b200.utils import principal_angle
    lead = edr.components_[0] / np.linalg.norm(edr.components_[0]); Bt = B.cpu().numpy().T
    print(json.dumps({'input': name, 'ms': min(ts) * 1e3, 'pts_per_s': n / min(ts), 'num_iter': edr.num_iter,
                      'components': list(edr.components_.shape),
                      'lead_angle_to_truth': float(np.arcsin(min(1, np.linalg.norm(lead - Bt.T.dot(Bt.dot(lead))))))}))
