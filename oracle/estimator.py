"""sklearn-style shells around the oracle models.  TEST ORACLE (see ``oracle/__init__.py``).

Mirrors ``edrgp/gp_model/base.py:46-257`` (``_BaseGP``) and ``edrgp/gp_model/regression.py:
10-157`` with ``oracle.gpy_restatement`` standing in for GPy, so the UNMODIFIED reference
orchestration (``edrgp.edr.EffectiveDimensionalityReduction``) can run on top of it.
"""
import pickle
from copy import deepcopy

import numpy as np
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.utils import check_X_y, check_array, assert_all_finite
from sklearn.utils.validation import check_is_fitted

from . import gpy_restatement as gpy

_KERNELS = {'RBF': gpy.RBF}


class _BaseGP(BaseEstimator):
    def fit(self, X, y, **opt_kws):                       # edrgp/gp_model/base.py:46-70
        X, y = self._check_data(X, y)
        self.n_features_ = X.shape[1]
        kernel = self._make_kernel()
        self.estimator_ = self._get_model(X, y, kernel)
        opt_kws.setdefault('messages', False)
        opt_kws.setdefault('max_iters', 1000)
        getattr(self.estimator_, self.method)(**opt_kws)
        return self

    def _check_data(self, X, y):                          # :72-91
        X, y = check_X_y(X, y, accept_sparse=False)
        return X, y[:, np.newaxis]

    def _check_input(self, X):                            # :93-109
        X = check_array(X, accept_sparse=False)
        if X.shape[1] != self.n_features_:
            raise ValueError("X has {} features per sample; expecting {}"
                             .format(X.shape[1], self.n_features_))
        return X

    def _make_kernel(self):                               # :111-147
        if self.kernels is None:
            return None
        if isinstance(self.kernels, gpy.RBF):
            return self.kernels.copy()
        kernels = [self.kernels] if isinstance(self.kernels, str) else list(self.kernels)
        options = self.kernel_options
        if isinstance(options, dict):
            options = [options]
        input_dim = {'input_dim': self.n_features_}
        if options is None:
            options = [dict(input_dim) for _ in kernels]
        elif len(kernels) == len(options):
            options = deepcopy(options)
            for opt in options:
                opt.update(input_dim)
        else:
            raise ValueError
        if len(kernels) != 1:
            raise NotImplementedError("sums of kernels are outside the restated path")
        return _KERNELS[kernels[0]](**options[0])

    def _check_predict(self, X):
        X = self._check_input(X)
        check_is_fitted(self, 'estimator_')
        return X

    def predict(self, X):                                 # :169-189
        X = self._check_predict(X)
        y_pred = self.estimator_.predict(X)[0][:, 0]
        assert_all_finite(y_pred)
        return y_pred

    def predict_variance(self, X):                        # :191-206
        X = self._check_predict(X)
        return self.estimator_.predict(X)[1]

    def predict_gradient(self, X):                        # :208-222
        X = self._check_predict(X)
        return self.estimator_.predictive_gradients(X)[0][:, :, 0]

    def save(self, model_path):                           # :224-239
        if not model_path.endswith('.pickle'):
            model_path += '.pickle'
        with open(model_path, 'wb') as f:
            pickle.dump(self.estimator_, f)

    def load(self, model_path):                           # :242-257
        if not model_path.endswith('.pickle'):
            model_path += '.pickle'
        with open(model_path, 'rb') as f:
            self.estimator_ = pickle.load(f)


class GaussianProcessRegressor(_BaseGP, RegressorMixin):
    """Dense GP (edrgp/gp_model/regression.py:10-77); used only to pin test_sparse_regression."""

    def __init__(self, kernels=None, kernel_options=None, Y_metadata=None, normalizer=True,
                 noise_var=1.0, mean_function=None, method='optimize'):
        self.normalizer = normalizer
        self.noise_var = noise_var
        self.kernels = kernels
        self.kernel_options = kernel_options
        self.Y_metadata = Y_metadata
        self.mean_function = mean_function
        self.method = method

    def _get_model(self, X, y, kernel):
        return gpy.GPRegression(X, y, kernel, self.normalizer, self.noise_var)


class SparseGaussianProcessRegressor(_BaseGP, RegressorMixin):
    """Sparse GP (edrgp/gp_model/regression.py:80-157)."""

    def __init__(self, kernels=None, kernel_options=None, Z=None, num_inducing=10,
                 Y_metadata=None, X_variance=None, normalizer=True, mean_function=None,
                 method='optimize'):
        self.kernels = kernels
        self.kernel_options = kernel_options
        self.Z = Z
        self.num_inducing = num_inducing
        self.Y_metadata = Y_metadata
        self.X_variance = X_variance
        self.normalizer = normalizer
        self.mean_function = mean_function
        self.method = method

    def _get_model(self, X, y, kernel):
        return gpy.SparseGPRegression(X, y, kernel=kernel, Z=self.Z,
                                      num_inducing=self.num_inducing,
                                      X_variance=self.X_variance,
                                      mean_function=self.mean_function,
                                      normalizer=self.normalizer)
