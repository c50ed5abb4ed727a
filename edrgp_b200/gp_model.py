"""sklearn-style estimator shell around the device model: drop-in for edr-gp's
``SparseGaussianProcessRegressor`` (``edrgp/gp_model/regression.py:80-157``) with the methods of
its ``_BaseGP`` (``edrgp/gp_model/base.py:46-257``): ``fit``, ``predict``, ``predict_variance``,
``predict_gradient``, ``save``, ``load`` -- same constructor parameters, argument meaning and
errors.  The GPy model behind it is replaced by ``edrgp_b200.model.SparseGPRegression``.

Only what the hot path needs is provided: the RBF kernel (ARD or not) with a Gaussian likelihood.
Other GPy kernels, sums of kernels, uncertain inputs and mean functions raise
``NotImplementedError`` instead of silently doing something else.
"""
from copy import deepcopy

import numpy as np
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.utils import check_X_y, check_array, assert_all_finite
from sklearn.utils.validation import check_is_fitted

from . import model as _model

_KERNELS = {'RBF': _model.RBF, 'rbf': _model.RBF}


def _from_gpy_kernel(kern, n_features):
    """A ready ``GPy.kern`` object (edrgp/gp_model/base.py:124-126 hands it to the model as it is): an RBF
    over all the input columns is read into the B200 path's own hyper-parameter holder, everything else
    (sums, products, other covariance functions, ``active_dims`` subsets) is outside the path and says so.
    Duck-typed: GPy itself is never imported."""
    if getattr(kern, 'parts', None):
        raise NotImplementedError("sums / products of kernels are outside the B200 path (RBF only)")
    if getattr(kern, 'name', type(kern).__name__).lower() != 'rbf' and type(kern).__name__ != 'RBF':
        raise NotImplementedError("kernel %r is outside the B200 path (RBF only)" % (type(kern).__name__,))
    input_dim = int(kern.input_dim)
    if input_dim != int(n_features):
        raise ValueError("kernel has input_dim {}; X has {} features per sample".format(input_dim, n_features))
    dims = getattr(kern, 'active_dims', None)
    if dims is not None and not np.array_equal(np.asarray(dims).ravel(), np.arange(input_dim)):
        raise NotImplementedError("active_dims subsets are outside the B200 path")
    variance = float(np.asarray(kern.variance, dtype=np.float64).ravel()[0])
    lengthscale = np.asarray(kern.lengthscale, dtype=np.float64).ravel()
    return _model.RBF(input_dim, variance, lengthscale, ARD=bool(getattr(kern, 'ARD', lengthscale.size > 1)))


def _host_copy_threads():
    """Host threads that stage pageable rows into pinned blocks: the CPUs this process may use, shared fairly
    between the ranks of this node, one left for the thread that drives the GPU; at most 8 (memcpy saturates the
    memory channels well before that)."""
    import os
    try:
        cpus = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cpus = os.cpu_count() or 1
    local = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    env = os.environ.get('EDRGP_H2D_THREADS')
    if env:
        return max(1, int(env))
    return max(1, min(8, cpus // local - 1))


class _BaseGP(BaseEstimator):
    """Common estimator logic (mirrors ``edrgp/gp_model/base.py:_BaseGP``)."""

    def fit(self, X, y, **opt_kws):
        """Fit the model: build it on the device, then run ``method`` (``'optimize'``,
        ``'optimize_restarts'`` or ``'fixed'``) with ``messages=False, max_iters=1000`` defaults
        (edrgp/gp_model/base.py:46-70)."""
        X, y = self._check_data(X, y)
        previous = {k: getattr(self, k) for k in ('n_features_', 'estimator_') if hasattr(self, k)}
        try:
            self.n_features_ = X.shape[1]
            kernel = self._make_kernel()
            model = self._get_model(X, y, kernel)
            opt_kws.setdefault('messages', False)
            opt_kws.setdefault('max_iters', 1000)
            getattr(model, self.method)(**opt_kws)
            if not getattr(self, 'deferred_checks', False):
                model._check_pd()      # deferred input / positive-definiteness checks surface here
        except Exception:
            # a failed fit leaves the estimator as it was (the reference validates before it builds)
            for k in ('n_features_', 'estimator_'):
                if k in previous:
                    setattr(self, k, previous[k])
                elif hasattr(self, k):
                    delattr(self, k)
            raise
        self.estimator_ = model
        return self

    def _check_data(self, X, y):
        X, y = check_X_y(X, y, accept_sparse=False, dtype=np.float64)
        return X, y[:, np.newaxis]

    def _check_input(self, X):
        X = check_array(X, accept_sparse=False, dtype=np.float64)
        if X.shape[1] != self.n_features_:
            raise ValueError("X has {} features per sample; expecting {}".format(X.shape[1], self.n_features_))
        return X

    def _make_kernel(self):
        """Kernel from names + options (edrgp/gp_model/base.py:111-147).  ``None`` lets the model
        pick its default (RBF, as GPy does); a ready ``edrgp_b200.model.RBF`` passes through."""
        if self.kernels is None:
            return None
        if isinstance(self.kernels, _model.RBF):
            return self.kernels.copy()
        if (getattr(self.kernels, '__module__', None) or '').startswith('GPy.kern'):
            return _from_gpy_kernel(self.kernels, self.n_features_)
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
            raise ValueError("kernels and kernel_options differ in length")
        if len(kernels) != 1:
            raise NotImplementedError("sums of kernels are outside the B200 path (RBF only)")
        if kernels[0] not in _KERNELS:
            raise NotImplementedError("kernel %r is outside the B200 path (RBF only)" % (kernels[0],))
        return _KERNELS[kernels[0]](**options[0])

    def _check_predict(self, X):
        X = self._check_input(X)
        check_is_fitted(self, 'estimator_')
        return X

    def predict(self, X):
        """Posterior mean, shape (n,) (edrgp/gp_model/base.py:169-189)."""
        X = self._check_predict(X)
        y_pred = self.estimator_.predict(X, want_variance=False)[0][:, 0]
        assert_all_finite(y_pred)
        return y_pred

    def predict_variance(self, X):
        """Predictive variance, shape (n, 1) (edrgp/gp_model/base.py:191-206)."""
        X = self._check_predict(X)
        return self.estimator_.predict(X)[1]

    def predict_gradient(self, X):
        """Gradient of the posterior mean, shape (n, d) (edrgp/gp_model/base.py:208-222)."""
        X = self._check_predict(X)
        return self.estimator_.predictive_gradients(X)[0][:, :, 0]

    def save(self, model_path):
        """Save the fitted model to ``model_path`` (+ '.pickle', the reference's file name convention,
        edrgp/gp_model/base.py:224-239).  The CONTENT is not a GPy pickle: it is a NumPy ``.npz`` archive of plain
        arrays (hyper-parameters, Z, alpha, Cholesky factors, normaliser moments) that loads without executing
        anything (``allow_pickle=False``); files written by the reference cannot be read and vice versa."""
        check_is_fitted(self, 'estimator_')
        if not model_path.endswith('.pickle'):
            model_path += '.pickle'
        state = self.estimator_.state_dict()
        arrays = {'format': np.array('edrgp_b200/2'), 'n_features_': np.array(self.n_features_)}
        arrays.update({'state_' + k: np.asarray(v) for k, v in state.items()})
        with open(model_path, 'wb') as f:
            np.savez(f, **arrays)

    def load(self, model_path):
        """Load a model saved by ``save`` (edrgp/gp_model/base.py:242-257).  The restored estimator predicts
        (``predict``, ``predict_variance``, ``predict_gradient``); it holds no training rows."""
        if not model_path.endswith('.pickle'):
            model_path += '.pickle'
        try:
            with np.load(model_path, allow_pickle=False) as blob:
                if 'format' not in blob.files or str(blob['format']) != 'edrgp_b200/2':
                    raise ValueError("not an edrgp_b200 model file")
                state = {k[6:]: blob[k] for k in blob.files if k.startswith('state_')}
                n_features = int(blob['n_features_'])
        except (OSError, ValueError) as e:
            raise ValueError("not an edrgp_b200 model file (%s); GPy pickles written by the reference are not "
                             "supported" % (e,))
        for k in ('variance', 'noise_variance', 'log_likelihood', 'y_mean', 'y_std'):
            if k in state:
                state[k] = float(state[k])
        for k in ('input_dim', 'num_data'):
            state[k] = int(state[k])
        state['ARD'] = bool(state['ARD'])
        self.estimator_ = _model.FittedSparseGP(state)
        self.n_features_ = n_features


class SparseGaussianProcessRegressor(_BaseGP, RegressorMixin):
    """Sparse Gaussian process regression on B200 (drop-in for
    ``edrgp.gp_model.SparseGaussianProcessRegressor``, edrgp/gp_model/regression.py:80-157).

    Parameters are the reference's; ``method`` additionally accepts ``'fixed'`` (keep the initial
    or injected hyper-parameters: what the parity harness and the benchmark use), and
    ``chunk_rows`` bounds the rows whose cross-covariance block is in HBM at a time, and
    ``noise_var`` sets the initial Gaussian noise variance (GPy's default 1.0; the dense
    ``GaussianProcessRegressor`` of the reference has the same parameter).
    ``precision='tf32x3'`` runs the large contractions over the training rows on the TF32-split tcgen05
    kernels where they apply -- cross-covariance and posterior-mean gradients for at most 64 features,
    the weights of the hyper-parameter gradient for at most 512 inducing points -- with entries within
    1e-4 relative of the FP64 ones; the default ``'fp64'`` is the reference's arithmetic.
    ``deferred_checks=True`` keeps ``fit`` from synchronising with the device: the non-finite scan
    of the input (sklearn's ``check_X_y``) and the positive-definiteness flag of the Cholesky step
    are still computed, but they are read -- and raise -- at the first host read-back
    (``predict*``, ``estimator_.gradient_gram(check=True)`` or ``estimator_.finish_checks()``)
    instead of inside ``fit``.  With row shards of a few hundred thousand points per GPU that one
    round trip is a visible part of a sweep.
    """

    def __init__(self, kernels=None, kernel_options=None, Z=None, num_inducing=10, Y_metadata=None,
                 X_variance=None, normalizer=True, mean_function=None, method='optimize', chunk_rows=262144,
                 noise_var=1.0, deferred_checks=False, precision='fp64'):
        self.kernels = kernels
        self.kernel_options = kernel_options
        self.Z = Z
        self.num_inducing = num_inducing
        self.Y_metadata = Y_metadata
        self.X_variance = X_variance
        self.normalizer = normalizer
        self.mean_function = mean_function
        self.method = method
        self.chunk_rows = chunk_rows
        self.noise_var = noise_var
        self.deferred_checks = deferred_checks
        self.precision = precision

    def _get_model(self, X, y, kernel):
        import torch
        if self.Y_metadata is not None:
            raise NotImplementedError("Y_metadata is outside the B200 path (Gaussian likelihood only)")
        kw = dict(kernel=kernel, Z=self.Z, num_inducing=self.num_inducing, X_variance=self.X_variance,
                  mean_function=self.mean_function, normalizer=self.normalizer, chunk_rows=self.chunk_rows,
                  noise_var=self.noise_var, precision=getattr(self, 'precision', 'fp64'))
        from . import ops as _ops

        def nonfinite_check(Xc, yc, scan=True):
            """(device count of NaN / Inf entries, raiser): the model reads the count together with its
            other deferred scalars in one transfer and calls the raiser if it is non-zero.  scan=False: the
            caller has its own evidence (the cross-covariance kernel flags non-finite rows as it goes) and
            only wants the raiser."""
            bad = _ops.count_nonfinite(Xc, yc) if scan else None   # enqueued now, read at the first host sync

            def on_bad():
                bad_input = _model.NonFiniteInput         # a ValueError, as sklearn's validators raise
                if int(_ops.count_nonfinite(Xc).cpu()[0]) != 0:
                    raise bad_input("Input X contains NaN or infinity.")
                if int(_ops.count_nonfinite(yc).cpu()[0]) != 0:
                    raise bad_input("Input y contains NaN or infinity.")
                raise bad_input("Input contains NaN or infinity on another rank, or values too large for the "
                                "kernel arithmetic (|x / lengthscale|^2 overflows float64).")
            return bad, on_bad

        if isinstance(X, torch.Tensor):
            return _model.SparseGPRegression(X, y, pre_sync_check=lambda scan=True: nonfinite_check(X, y, scan), **kw)
        # Host rows: the library's streamer (edrgp_h2d_*) moves them block by block while the statistics pass already
        # works on the blocks that have arrived -- straight from the caller's buffer when it is pinned, through a ring
        # of pinned blocks filled by host threads when it is ordinary (pageable) memory; the non-finite scan
        # (sklearn's check_X_y) runs on the device, before the first host read-back.
        from . import _lib
        lib = _lib.load()
        n, d = X.shape
        dev = torch.device('cuda', torch.cuda.current_device())
        de = d + (d & 1)
        Xd = torch.zeros(n, de, dtype=torch.float64, device=dev) if de != d else \
            torch.empty(n, d, dtype=torch.float64, device=dev)
        yd = torch.empty(y.shape, dtype=torch.float64, device=dev)
        main = torch.cuda.current_stream(dev).cuda_stream
        # copy granularity: a quarter of chunk_rows (the size of the statistics blocks while the rows are
        # arriving, see SparseGPRegression._chunks); a pinned source is enqueued one chunk_rows ahead of its consumer
        rows = int(max(1024, min(self.chunk_rows // 4 if self.chunk_rows >= 65536 else self.chunk_rows, max(n, 1))))
        rows &= ~1
        ahead = int(max(rows, min(self.chunk_rows, max(n, 1))))
        threads = _host_copy_threads()
        state = {'x': None, 'closed': False}

        def start():
            # opened at the first request, i.e. AFTER the model has uploaded its hyper-parameters and inducing
            # inputs (the copy engine serves its queue in order); the targets travel on the same copy stream right
            # behind the first block of rows
            if state['x'] is None and not state['closed']:
                state['x'] = lib.edrgp_h2d_open(X.ctypes.data, Xd.data_ptr(), n, d * 8, de * 8, rows, threads, 3, main,
                                                y.ctypes.data, yd.data_ptr(), n * 8)
                if not state['x']:
                    raise _lib.EdrgpError("edrgp_h2d_open failed: %s" % lib.edrgp_last_error().decode())

        def loader(s, e):
            if state['closed']:
                return
            start()
            _lib.check(lib.edrgp_h2d_wait(state['x'], e, ahead, main), 'edrgp_h2d_wait')

        def handle():
            """The streamer's handle for the composite sweep calls, which wait block by block themselves."""
            start()
            return state['x']

        loader.handle, loader.rows, loader.ahead = handle, rows, ahead

        def y_loader():
            if state['closed']:
                return
            start()
            _lib.check(lib.edrgp_h2d_wait_side(state['x'], main), 'edrgp_h2d_wait_side')

        def close():
            if not state['closed']:
                state['closed'] = True
                if state['x']:
                    lib.edrgp_h2d_close(state['x'])           # joins the copy threads, waits for the DMA
                    state['x'] = None

        def check(scan=True):
            loader(0, n)
            y_loader()
            close()
            return nonfinite_check(Xd, yd, scan)

        try:
            return _model.SparseGPRegression(Xd, yd, input_dim=d, row_loader=loader, pre_sync_check=check,
                                             y_loader=y_loader, **kw)
        finally:
            close()                                  # (a no-op after a normal construction: check() has run)

    def _check_data(self, X, y):
        """Validation of ``check_X_y`` (edrgp/gp_model/base.py:72-91) split so that the O(n d) part
        runs where the data is going anyway: shapes and dtypes on the host, the non-finite scan on
        the device after the copy.  CUDA tensors (this rank's shard) are taken as they are."""
        import torch
        if isinstance(X, torch.Tensor):
            if X.dim() != 2 or y.shape[0] != X.shape[0]:
                raise ValueError("X must be (n, d) and y (n,)")
            Xd = X.to(device='cuda', dtype=torch.float64)
            yd = y.to(device='cuda', dtype=torch.float64).reshape(-1, 1)
            return Xd, yd                    # the non-finite scan runs on the device (see _get_model)
        X = np.asarray(X)
        y = np.asarray(y)
        if X.ndim != 2:
            raise ValueError("Expected 2D array, got %dD array instead" % X.ndim)
        if y.ndim == 2 and y.shape[1] == 1:
            y = y[:, 0]
        if y.ndim != 1:
            raise ValueError("y should be a 1d array, got an array of shape {} instead.".format(y.shape))
        if X.shape[0] != y.shape[0]:
            raise ValueError("Found input variables with inconsistent numbers of samples: [%d, %d]"
                             % (X.shape[0], y.shape[0]))
        if X.shape[0] < 1 or X.shape[1] < 1:
            raise ValueError("Found array with %d sample(s) and %d feature(s) while a minimum of 1 is required."
                             % X.shape)
        return np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)

    def _check_input(self, X):
        import torch
        if isinstance(X, torch.Tensor):
            if X.shape[1] != self.n_features_:
                raise ValueError("X has {} features per sample; expecting {}".format(X.shape[1], self.n_features_))
            return X
        return _BaseGP._check_input(self, X)
