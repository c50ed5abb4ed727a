"""Small-shape pass over every kernel family for compute-sanitizer (memcheck): ragged n, odd d / m,
feature-blocked d > 128, one optimisation step, prediction surface, EDR fit."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import edrgp_b200 as eb
from edrgp_b200 import model as emodel, ops

rng = np.random.RandomState(0)
for n, d, m in [(301, 7, 33), (257, 64, 40), (200, 130, 17), (129, 200, 65), (1, 4, 3)]:
    X = rng.standard_normal((n, d)); y = rng.standard_normal(n)
    Z = rng.standard_normal((m, d)); ell = 1.0 + rng.uniform(size=d) * np.sqrt(d)
    mod = emodel.SparseGPRegression(X, y[:, None], kernel=emodel.RBF(d, 1.3, ell, ARD=True), Z=Z, normalizer=True,
                                    chunk_rows=1024, noise_var=0.2)
    mod.log_likelihood()
    G, C = mod.gradient_gram(want_G=True, want_C=True)
    mod.predictive_gradients(X[: max(1, n // 2)])
    mod.predict(X[: max(1, n // 2)])
    if n > 1:
        mod._need_grad = True; mod.parameters_changed(); mod._need_grad = False
    eb.GramEighTransformer().fit_gram(C, n)
    torch.cuda.synchronize()
    print('ok', n, d, m)
X = rng.standard_normal((400, 6)); yv = np.tanh(X[:, 0]) + 0.1 * rng.standard_normal(400)
np.random.seed(0)
edr = eb.EffectiveDimensionalityReduction(eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=20),
                                          eb.GramEighTransformer(), n_components=1, step=2,
                                          preprocessor=eb.DevicePCA(n_components=5)).fit(X, yv, max_iters=3)
print('edr ok', edr.components_.shape)
torch.cuda.synchronize()
