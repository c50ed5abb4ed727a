"""Wall-clock (synchronised) phases of one fixed-hyper-parameter sweep (development aid)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops
n, d, m = 4_000_000, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = sync()
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed', noise_var=0.1, chunk_rows=524288)
    Xc, yc = est._check_data(X, y)
    t1 = sync()
    est.n_features_ = d
    mod = est._get_model(Xc, yc, est._make_kernel())
    t2 = sync()
    mod._check_pd()
    est.estimator_ = mod
    t3 = sync()
    _, C = mod.gradient_gram(want_G=False)
    t4 = sync()
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n)
    t5 = sync()
    print(json.dumps({'check_data': (t1 - t0) * 1e3, 'model_ctor': (t2 - t1) * 1e3, 'check_pd': (t3 - t2) * 1e3,
                      'gradient_gram': (t4 - t3) * 1e3, 'fit_gram': (t5 - t4) * 1e3, 'total': (t5 - t0) * 1e3}))
    del est, mod
# finer: inside the model constructor
import cProfile, pstats
est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True, method='fixed', noise_var=0.1, chunk_rows=524288)
pr = cProfile.Profile(); pr.enable()
est.fit(X, y); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
