"""CPU restatement of edr-gp's EDR orchestration loop.  TEST ORACLE (see ``oracle/__init__.py``).

``/root/reference`` does not exist on the GPU box, so the GPU parity tests cannot import the
reference's ``EffectiveDimensionalityReduction``.  This module restates its control flow
(``edrgp/edr.py:90-113,142-176,215-241,261-289`` on top of ``edrgp/base.py:115-140,172-200,
435-517``) in one function on top of the oracle estimator (``oracle/estimator.py``) and the
reference's SVD semantics (``edrgp/utils.py:27-55,123-157``, economy SVD: same Vh and S).
``tests/test_oracle.py::test_reference_loop_matches_unmodified_reference`` pins it, in the
container that has ``/root/reference``, against the UNMODIFIED reference classes run with the same
oracle estimator; ``tests/golden/`` holds outputs generated that way.
"""
import numpy as np

from .estimator import SparseGaussianProcessRegressor


def _subspace_variance_ratio(X, V):                      # edrgp/utils.py:27-55
    if np.allclose(np.dot(V.T, V), np.eye(V.shape[1])):
        var = np.linalg.norm(X.dot(V), axis=0)
    else:
        var = np.linalg.norm(X.dot(np.linalg.qr(V)[0]))
    return var, (var / np.linalg.norm(X)) ** 2


def _svd_components(G):                                  # edrgp/utils.py:123-157 with n_components=None
    _, S, Vh = np.linalg.svd(G, full_matrices=False)
    k = min(G.shape[0], G.shape[1])
    return Vh[:k], (S ** 2)[:k]


def fit_reference_style(X, y, num_inducing=10, n_components=None, step=None, normalize=True,
                        kernels='RBF', kernel_options=None, Z=None, method='optimize', **opt_kws):
    kernel_options = {'ARD': True} if kernel_options is None else kernel_options
    X = np.asarray(X, dtype=np.float64)

    def new_estimator():
        return SparseGaussianProcessRegressor(kernels, kernel_options, Z=Z, num_inducing=num_inducing,
                                              method=method)

    # EDR._preprocessing_fit (edrgp/edr.py:142-176), no preprocessor
    if normalize:
        mean, scale = X.mean(0), X.std(0)
        scale = np.where(scale < 10 * np.finfo(float).eps, 1.0, scale)
        Xp = (X - mean) / scale
        reverse_scaling = np.diag(1 / scale)
    else:
        Xp = X
    d = Xp.shape[1]
    k_target = d if n_components is None else n_components          # _check_init
    adaptive = False                                                 # _check_step
    if step is None:
        step_ = k_target
    elif isinstance(step, int) and step > 0:
        if k_target == d:
            raise ValueError("If step is int (n_components < n_features) must be True")
        step_ = step
    elif isinstance(step, float) and 0 < step < 1:
        if n_components is not None:
            raise ValueError("If step is float n_components should be None")
        adaptive, step_ = True, step
    else:
        raise ValueError("Step should be None or int > 0 or float from 0 to 1")

    components = None
    first_gradients = None
    num_iter = 0
    X_proj = Xp.copy()
    cont = True
    while cont:                                                      # IterativeEDR.fit (base.py:459-463)
        est = new_estimator().fit(X_proj, y, **opt_kws)
        grad = est.predict_gradient(X_proj)
        if num_iter == 0:
            first_gradients = grad
        comps, _ = _svd_components(grad)
        if adaptive:                                                 # _select_n_components
            _, vr = _subspace_variance_ratio(grad, comps.T)
            n_sel = int(np.sum(np.cumsum(vr) < step_, dtype=int)) + 1
            if n_sel == grad.shape[1]:
                cont = False
        else:
            n_sel = max(k_target, grad.shape[1] - step_)
            if n_sel == k_target:
                cont = False
        components = comps if components is None else np.dot(comps, components)   # _select_best_components
        _, vr = _subspace_variance_ratio(first_gradients, components.T)
        best = np.argsort(vr)[-n_sel:][::-1]
        components = components[best, :]
        X_proj = np.dot(Xp, components.T)
        num_iter += 1
    est = new_estimator().fit(X_proj, y, **opt_kws)                   # _last_fit (base.py:172-200)
    subspace_gradients = est.predict_gradient(X_proj)
    var, ratio = _subspace_variance_ratio(first_gradients, components.T)
    out_components = np.dot(components, reverse_scaling) if normalize else components
    return {'components_': out_components, 'num_iter': num_iter, '_first_gradients_': first_gradients,
            'subspace_gradients_': subspace_gradients, 'subspace_variance_': var,
            'subspace_variance_ratio_': ratio, 'estimator_': est}


class EconomySVDTransformer(object):
    """Host transformer with the semantics of the reference's ``SVDTransformer`` (edrgp/utils.py:81-175:
    uncentred PCA, ``components_ = Vh[:k]``, ``subspace_variance_ = S^2``, ratio ``S^2 / sum S^2``, k capped
    by the number of rows) on the ECONOMY SVD -- same Vh and S, without the n x n ``U`` that makes the
    reference's own class unusable beyond a few ten thousand rows.  TEST ORACLE: the GPU tests use it as "any
    host transformer with fit + components_" where the reference copy is not needed."""

    def __init__(self, n_components=None):
        self.n_components = n_components

    def get_params(self, deep=True):
        return {'n_components': self.n_components}

    def set_params(self, **params):
        for k, v in params.items():
            setattr(self, k, v)
        return self

    def fit(self, X, y=None):
        X = np.array(X, dtype=np.float64)
        _, S, Vh = np.linalg.svd(X, full_matrices=False)
        ratio = S ** 2 / np.sum(S ** 2)
        nc, k = self.n_components, X.shape[1]
        if isinstance(nc, (int, np.integer)) and 0 < nc <= X.shape[1]:
            k = int(nc)
        elif isinstance(nc, float) and 0 < nc < 1:
            k = int(np.sum(np.cumsum(ratio) < nc, dtype=int)) + 1
        k = min(X.shape[0], k)
        self.components_ = Vh[:k, :]
        self.subspace_variance_ = (S ** 2)[:k]
        self.subspace_variance_ratio_ = ratio[:k]
        return self

    def transform(self, X):
        return np.asarray(X).dot(self.components_.T)
