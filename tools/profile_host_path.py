"""Host-side (Python) cost of one sweep through the public classes on a small shard, where the device work is short
enough that the host is what the device waits for: cProfile over many sweeps + time to the first kernel launch."""
import cProfile
import json
import pstats
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops

n, d, m = 65536, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].cpu().numpy()
ell = np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d))
marks = {}
orig_begin = ops.FixedSweep.begin
orig_eigh = ops.FixedSweep.eigh


def begin(self, *a, **k):
    marks['begin'] = time.perf_counter()
    return orig_begin(self, *a, **k)


def eigh(self, *a):
    marks['eigh_in'] = time.perf_counter()
    r = orig_eigh(self, *a)
    marks['eigh_out'] = time.perf_counter()
    return r


ops.FixedSweep.begin = begin
ops.FixedSweep.eigh = eigh


def sweep():
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True,
                                            method='fixed', noise_var=0.1, chunk_rows=524288, deferred_checks=True).fit(X, y)
    marks['fit_done'] = time.perf_counter()
    _, C = est.estimator_.gradient_gram(want_G=False, check=False)
    marks['grad_done'] = time.perf_counter()
    tr = eb.GramEighTransformer(n_components=3).fit_gram(C, n)
    est.estimator_.finish_checks()
    return tr.components_


for _ in range(20):
    sweep()
acc = {}
reps = 200
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sweep()
    t1 = time.perf_counter()
    for k, v in (('to_first_launch', marks['begin'] - t0), ('fit_rest', marks['fit_done'] - marks['begin']),
                 ('gradient_gram', marks['grad_done'] - marks['fit_done']), ('to_eigh_call', marks['eigh_in'] - marks['grad_done']),
                 ('eigh_call_incl_wait', marks['eigh_out'] - marks['eigh_in']), ('after_readback', t1 - marks['eigh_out']),
                 ('total', t1 - t0)):
        acc[k] = acc.get(k, 0.0) + v * 1e3 / reps
print(json.dumps({'host_ms': acc}))
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    sweep()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
