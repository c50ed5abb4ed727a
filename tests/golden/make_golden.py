"""Generates the golden fixtures of tests/golden/ in the container that has /root/reference.

Each case runs the UNMODIFIED reference orchestration (edrgp.edr.EffectiveDimensionalityReduction
with edrgp.utils.SVDTransformer, imported from /root/reference) on top of the oracle estimator
(oracle/estimator.py: the NumPy restatement of GPy, which is not installable here) and stores the
inputs and the fitted attributes.  The reference's GP arithmetic itself (GPy) cannot run, so these
vectors pin (a) the oracle's orchestration against the real reference L3 and (b) the CUDA path
against the oracle chain; see oracle/__init__.py ("parity unpinned" at the GPy boundary).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.append('/root/reference')

from edrgp.edr import EffectiveDimensionalityReduction      # noqa: E402  (the reference, unmodified)
from edrgp.utils import SVDTransformer                      # noqa: E402
from edrgp.datasets import get_beta_inputs, get_edr_target  # noqa: E402
from oracle.estimator import SparseGaussianProcessRegressor  # noqa: E402
from oracle import gpy_restatement as gpy                    # noqa: E402

CASES = {
    # BASELINE config 1: n=500, d=10, m=20 (the BriefIntro recipe, examples/BriefIntro.ipynb:657-662)
    'c1_onepass': dict(n=500, d=10, m=20, k=2, step=None, normalize=True, max_iters=0, seed=3),
    'c1_iterative': dict(n=500, d=10, m=20, k=2, step=3, normalize=True, max_iters=0, seed=3),
    'small_adaptive': dict(n=300, d=6, m=15, k=None, step=0.97, normalize=False, max_iters=0, seed=5),
    'small_optimised': dict(n=300, d=5, m=15, k=1, step=None, normalize=True, max_iters=25, seed=8),
}


def make_data(n, d, seed):
    np.random.seed(seed)
    X = get_beta_inputs(n, d)
    B = np.linalg.qr(np.random.randn(d, 2))[0]
    y = get_edr_target(X.dot(B), 0.1)
    return X, np.asarray(y, dtype=float), B


def main():
    for name, c in CASES.items():
        X, y, B = make_data(c['n'], c['d'], c['seed'])
        np.random.seed(100 + c['seed'])
        edr = EffectiveDimensionalityReduction(
            SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=c['m']), SVDTransformer(),
            n_components=c['k'], step=c['step'], normalize=c['normalize'])
        edr.fit(X, y, max_iters=c['max_iters'])
        mod = edr.estimator_.estimator_
        # a fixed-hyper-parameter snapshot of the FIRST model for kernel-level parity
        np.random.seed(100 + c['seed'])
        Xs = (X - X.mean(0)) / X.std(0) if c['normalize'] else X
        first = gpy.SparseGPRegression(Xs, y[:, None], kernel=gpy.RBF(c['d'], ARD=True), num_inducing=c['m'],
                                       normalizer=True)
        out = dict(X=X, y=y, B=B,
                   components_=edr.components_, num_iter=edr.num_iter,
                   subspace_variance_=edr.subspace_variance_, subspace_variance_ratio_=edr.subspace_variance_ratio_,
                   first_gradients=edr._first_gradients_, subspace_gradients_=edr.subspace_gradients_,
                   final_loglik=float(mod.log_likelihood()[0, 0]),
                   first_Z=first.Z, first_alpha=first.posterior.woodbury_vector[:, 0],
                   first_loglik=float(first.log_likelihood()[0, 0]),
                   first_grad_Z=first.grad_Z, first_grad_lengthscale=first.grad_lengthscale,
                   first_grad_variance=first.grad_variance, first_grad_noise=first.grad_noise,
                   first_pred_grad=first.predictive_gradients(Xs[:64])[0][:, :, 0],
                   first_pred_mean=first.predict(Xs[:64])[0][:, 0], first_pred_var=first.predict(Xs[:64])[1][:, 0])
        # the public surface behind the fit (edrgp/base.py:202-239, edrgp/edr.py:115-140,199-289): refit on the
        # kept first gradients (all rows, a row subset, a sparse transformer), gradients of the final
        # estimator at new rows mapped back to the raw features, projection, feature importances
        kk = 2 if c['k'] is None else c['k']
        Xq = X[:40] + 0.01
        out.update(get_estimator_gradients=edr.get_estimator_gradients(Xq), transform=edr.transform(Xq),
                   feature_importances_=edr.feature_importances_)
        edr.refit(SVDTransformer(n_components=kk))
        out.update(refit_components_=edr.refit_components_, refit_subspace_variance_=edr.refit_subspace_variance_,
                   refit_subspace_variance_ratio_=edr.refit_subspace_variance_ratio_,
                   refit_transform=edr.transform(Xq, refitted=True))
        rows = np.arange(0, c['n'], 3)
        edr.refit(SVDTransformer(n_components=kk), rows)
        out.update(refit_rows=rows, refit_rows_components_=edr.refit_components_,
                   refit_rows_subspace_variance_ratio_=edr.refit_subspace_variance_ratio_)
        from sklearn.decomposition import SparsePCA
        edr.refit(SparsePCA(n_components=kk, alpha=0.5, random_state=0))
        out.update(refit_sparse_components_=edr.refit_components_,
                   refit_sparse_subspace_variance_ratio_=edr.refit_subspace_variance_ratio_)
        for k, v in c.items():
            out['cfg_' + k] = np.array(-1 if v is None else v)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, edr.num_iter, edr.components_.shape, out['final_loglik'])


if __name__ == '__main__':
    main()
