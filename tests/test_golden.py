"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the UNMODIFIED
reference orchestration + the oracle estimator): the CPU oracle chain must reproduce them (here,
no GPU), and the CUDA classes must match them within the FP64 tolerances (``-m gpu``)."""
import glob
import os

import numpy as np
import pytest

from oracle import gpy_restatement as gpy
from oracle import pipeline as op
from oracle import reference_loop as oloop

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz')))


def _cfg(g):
    c = {k[4:]: g[k].item() for k in g.files if k.startswith('cfg_')}
    for k in ('k', 'step'):
        if c[k] == -1:
            c[k] = None
    if c['step'] is not None and float(c['step']) >= 1:
        c['step'] = int(c['step'])
    if c['k'] is not None:
        c['k'] = int(c['k'])
    return c


def test_fixtures_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_chain_reproduces_golden(path):
    g = np.load(path)
    c = _cfg(g)
    np.random.seed(100 + int(c['seed']))
    out = oloop.fit_reference_style(g['X'], g['y'], num_inducing=int(c['m']), n_components=c['k'], step=c['step'],
                                    normalize=bool(c['normalize']), max_iters=int(c['max_iters']))
    assert out['num_iter'] == int(g['num_iter'])
    assert np.allclose(out['components_'], g['components_'], rtol=1e-9, atol=1e-12)
    assert np.allclose(out['subspace_variance_ratio_'], g['subspace_variance_ratio_'], rtol=1e-9)
    assert np.allclose(out['_first_gradients_'], g['first_gradients'], rtol=1e-9, atol=1e-12)
    assert abs(float(out['estimator_'].estimator_.log_likelihood()[0, 0]) - float(g['final_loglik'])) \
        < 1e-9 * abs(float(g['final_loglik']))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_model_matches_golden_first_model(path):
    """Kernel-level parity at the initial hyper-parameters: bound, posterior mean / variance /
    gradients, and the hyper-parameter gradients of the first model of the run."""
    from edrgp_b200 import model
    g = np.load(path)
    c = _cfg(g)
    X, y = g['X'], g['y']
    Xs = (X - X.mean(0)) / X.std(0) if c['normalize'] else X
    mod = model.SparseGPRegression(Xs, y[:, None], kernel=model.RBF(X.shape[1], ARD=True), Z=g['first_Z'],
                                   normalizer=True)
    ll = float(mod.log_likelihood()[0, 0])
    assert abs(ll - float(g['first_loglik'])) < 1e-9 * abs(float(g['first_loglik']))

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

    assert rel(mod.predictive_gradients(Xs[:64])[0][:, :, 0], g['first_pred_grad']) < 1e-8
    mu, var = mod.predict(Xs[:64])
    assert rel(mu[:, 0], g['first_pred_mean']) < 1e-8
    assert rel(var[:, 0], g['first_pred_var']) < 1e-8
    mod._need_grad = True
    mod.parameters_changed()
    assert rel(mod.grad_Z, g['first_grad_Z']) < 1e-7
    assert rel(mod.grad_lengthscale, g['first_grad_lengthscale']) < 1e-7
    assert abs(mod.grad_variance - float(g['first_grad_variance'])) < 1e-7 * max(1, abs(float(g['first_grad_variance'])))
    assert abs(mod.grad_noise - float(g['first_grad_noise'])) < 1e-7 * max(1, abs(float(g['first_grad_noise'])))


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in GOLDEN if 'optimised' not in p],
                         ids=[os.path.basename(p)[:-4] for p in GOLDEN if 'optimised' not in p])
def test_cuda_edr_matches_golden(path):
    """The whole EDR fit at fixed hyper-parameters against the reference-orchestrated fixture."""
    import edrgp_b200 as eb
    g = np.load(path)
    c = _cfg(g)
    np.random.seed(100 + int(c['seed']))
    edr = eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=int(c['m'])), eb.GramEighTransformer(),
        n_components=c['k'], step=c['step'], normalize=bool(c['normalize']))
    edr.fit(g['X'], g['y'], max_iters=int(c['max_iters']))
    assert edr.num_iter == int(g['num_iter'])
    assert op.principal_angle(edr.components_, g['components_']) < 1e-6
    assert np.allclose(edr.subspace_variance_ratio_, g['subspace_variance_ratio_'], rtol=1e-7, atol=1e-12)
    fg = g['first_gradients']
    assert np.max(np.abs(edr._first_gradients_ - fg)) < 1e-8 * np.max(np.abs(fg))
    ll = float(edr.estimator_.estimator_.log_likelihood()[0, 0])
    assert abs(ll - float(g['final_loglik'])) < 1e-8 * abs(float(g['final_loglik']))
