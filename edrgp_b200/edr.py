"""``EffectiveDimensionalityReduction`` for the B200 path: same constructor, methods and fitted
attributes as edr-gp's (``edrgp/edr.py:11-289`` on top of ``edrgp/base.py:13-517``), with the
n-scale arithmetic kept on the device:

* the rows live in HBM for the whole fit (scaled once by ``edrgp_col_moments`` /
  ``edrgp_standardize``; projected between iterations by ``edrgp_project``);
* the estimator is fitted on device rows and its posterior-mean gradients are reduced to the d x d
  Gram matrix C = G^T G by the fused kernel; the transformer sees C (``fit_gram``) and every
  subspace-variance ratio of the iteration logic is evaluated from the first iteration's C
  (||G v||^2 = v^T C v, ||G||_F^2 = tr C) instead of another pass over G;
* with several processes (one per GPU, ``torch.distributed`` initialised) every rank passes its own
  rows; only 2d moments and d x d / m x m partial sums cross NVLink, and every rank ends with the
  same ``components_``.

The iteration logic -- ``step`` semantics, component selection against the *first* gradients, the
extra estimator fit on the projected data, reverse scaling of the components -- follows the
reference line by line in behaviour (see the citations on each method), not in code.
"""
import warnings
from copy import deepcopy

import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin, clone
from sklearn.preprocessing import StandardScaler, normalize as _l2_normalize
from sklearn.utils import check_array
from sklearn.utils.validation import check_is_fitted

from . import dist, ops
from .utils import subspace_variance_ratio_from_gram

F64 = torch.float64


def _to_device_rows(X):
    """Rows on the device.  Host arrays get sklearn's check_array semantics with the O(n d) part on
    the device: shape / dtype checks here, one copy, the NaN / Inf scan as a device pass."""
    if isinstance(X, torch.Tensor):
        Xd = X.to(dtype=F64)
        Xd = (Xd if Xd.is_cuda else Xd.cuda()).contiguous()
    else:
        X = np.asarray(X)
        if X.ndim != 2:
            raise ValueError("Expected 2D array, got %dD array instead" % X.ndim)
        if X.shape[0] < 1 or X.shape[1] < 1:
            raise ValueError("Found array with %d sample(s) and %d feature(s) while a minimum of 1 is required." % X.shape)
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to('cuda', non_blocking=True)
    if Xd.dim() != 2:
        raise ValueError("Expected 2D rows")
    if Xd.numel() and int(ops.count_nonfinite(Xd).cpu()[0]) != 0:
        raise ValueError("Input X contains NaN or infinity.")
    return Xd


def _to_device_targets(y, n):
    if isinstance(y, torch.Tensor):
        yd = y.to(dtype=F64).reshape(-1)
        yd = yd if yd.is_cuda else yd.cuda()
    else:
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        yd = torch.from_numpy(np.ascontiguousarray(y)).to('cuda', non_blocking=True)
    if yd.shape[0] != n:
        raise ValueError("Found input variables with inconsistent numbers of samples: [%d, %d]" % (n, yd.shape[0]))
    yd = yd.contiguous()
    if yd.numel() and int(ops.count_nonfinite(yd).cpu()[0]) != 0:
        raise ValueError("Input y contains NaN or infinity.")
    return yd


class EffectiveDimensionalityReduction(BaseEstimator, TransformerMixin):
    """Effective dimensionality reduction from posterior-mean gradients of a sparse GP.

    Parameters (as ``edrgp.EffectiveDimensionalityReduction``, edrgp/edr.py:83-88)
    ----------
    estimator : estimator whose fitted ``estimator_`` exposes ``gradient_gram`` (i.e.
        ``edrgp_b200.SparseGaussianProcessRegressor``).
    dr_transformer : transformer with ``fit`` + ``components_``; with ``fit_gram``
        (``GramEighTransformer``) the gradients never leave the device.
    n_components, step, normalize, preprocessor : as the reference.
    keep_gradients : bool (default True)
        Keep the first iteration's gradients on the device so that ``refit`` and
        ``_first_gradients_`` work (n x d doubles of HBM); the fit itself does not need them.
    """

    def __init__(self, estimator=None, dr_transformer=None, n_components=None, step=None, normalize=True,
                 preprocessor=None, keep_gradients=True):
        self.estimator = estimator
        self.dr_transformer = dr_transformer
        self.n_components = n_components
        self.step = step
        self.normalize = normalize
        self.preprocessor = preprocessor
        self.keep_gradients = keep_gradients

    @property
    def transformer(self):                      # the reference's attribute name (edrgp/base.py:409-414)
        return self.dr_transformer

    # ------------------------------------------------------------------------------------------
    # checks (edrgp/base.py:77-87, 416-433)
    # ------------------------------------------------------------------------------------------
    def _check_init(self, n_features):
        if self.estimator is None:
            raise ValueError("Estimator should be speciified")
        if self.dr_transformer is None:
            raise ValueError("transformer should be specified")
        self.n_components_ = n_features if self.n_components is None else self.n_components

    def _check_step(self, n_features):
        self.adaptive_step = False
        if self.step is None:
            self.step_ = self.n_components_
        elif isinstance(self.step, (int, np.integer)) and not isinstance(self.step, bool) and self.step > 0:
            if self.n_components_ == n_features:
                raise ValueError("If step is int (n_components < n_features) must be True")
            self.step_ = int(self.step)
        elif isinstance(self.step, float) and 0 < self.step < 1:
            if self.n_components is not None:
                raise ValueError("If step is float n_components should be None")
            self.adaptive_step = True
            self.step_ = self.step
        else:
            raise ValueError("Step should be None or int > 0 or float from 0 to 1")

    def _check_transformer(self, transformer):
        if not hasattr(transformer, 'components_'):
            raise AttributeError('The transformer does not expose "components_" attribute')

    # ------------------------------------------------------------------------------------------
    # preprocessing (edrgp/edr.py:142-197)
    # ------------------------------------------------------------------------------------------
    def _preprocessing_fit(self, Xd):
        """StandardScaler (population std, zero-variance columns left unscaled) on the device, then
        the optional linear preprocessor.  Returns the preprocessed device rows."""
        if not self.normalize:
            if self.preprocessor is not None:
                raise ValueError('To apply prerpocessing, normalize should be True')
            return Xd
        n_local, d = Xd.shape
        cnt = torch.tensor([float(n_local)], dtype=F64, device=Xd.device)
        if n_local:
            s1, _ = ops.col_moments(Xd)
            s1 = s1.clone()
        else:
            s1 = torch.zeros(d, dtype=F64, device=Xd.device)
        dist.allreduce_sum_(s1, cnt)
        n = float(cnt[0])
        mean = s1 / n
        if n_local:
            _, s2 = ops.col_moments(Xd, shift=mean)
            s2 = s2.clone()
        else:
            s2 = torch.zeros(d, dtype=F64, device=Xd.device)
        dist.allreduce_sum_(s2)
        var = s2 / n
        scale = torch.sqrt(var)
        # sklearn's _handle_zeros_in_scale: (near-)constant columns are not scaled
        eps = 10 * np.finfo(np.float64).eps
        scale = torch.where(scale < eps, torch.ones_like(scale), scale)
        scaler = StandardScaler()
        scaler.mean_ = mean.cpu().numpy()
        scaler.var_ = var.cpu().numpy()
        scaler.scale_ = scale.cpu().numpy()
        scaler.n_samples_seen_ = int(round(n))
        scaler.n_features_in_ = d
        self.scaler_ = scaler
        self._scaling_ = np.diag(scaler.scale_)
        self._reverse_scaling_ = np.diag(1 / scaler.scale_)
        Xs = ops.standardize(Xd, mean, scale) if n_local else Xd
        if self.preprocessor is not None:
            self.preprocessor_ = clone(self.preprocessor)
            if hasattr(self.preprocessor_, 'fit_transform_device'):
                # device preprocessor (DevicePCA): rows stay in HBM, moments reduced over ranks
                Xs = self.preprocessor_.fit_transform_device(Xs)
            else:
                if dist.is_distributed():
                    raise NotImplementedError("a host preprocessor needs all rows in one process; use DevicePCA")
                Xp = self.preprocessor_.fit_transform(Xs.cpu().numpy())
                Xs = torch.as_tensor(np.ascontiguousarray(Xp, dtype=np.float64), device=Xd.device)
            self._check_transformer(self.preprocessor_)
            self._preprocessing_ = self.preprocessor_.components_
        return Xs

    def _preprocessing_transform(self, X):
        X = check_array(X)
        if self.normalize is True:
            check_is_fitted(self, 'scaler_')
            X = self.scaler_.transform(X)
            X = np.dot(X, self._scaling_)
        return np.dot(X, self.components_.T)

    # ------------------------------------------------------------------------------------------
    # fit (edrgp/edr.py:90-113 around edrgp/base.py:435-466)
    # ------------------------------------------------------------------------------------------
    def fit(self, X, y=None, **opt_kws):
        """Fit on rows X (n, d) and targets y (n,): host arrays, or CUDA tensors holding this rank's
        shard.  ``opt_kws`` go to the estimator's ``fit`` unchanged."""
        if y is None:
            raise ValueError("y is required to fit the estimator")
        self.fitted = False
        Xd = _to_device_rows(X)
        yd = _to_device_targets(y, Xd.shape[0])
        self._n_local = Xd.shape[0]
        self._n_features_raw = Xd.shape[1]
        Xp = self._preprocessing_fit(Xd)
        if Xp is Xd and not isinstance(X, torch.Tensor):
            pass                                  # already a private device copy
        self._fit_iterations(Xp, yd, **opt_kws)
        if self.normalize:
            self.components_ = np.dot(self.components_, self._reverse_scaling_)
        self.fitted = True
        return self

    def _fit_iterations(self, Xp, yd, **opt_kws):
        n_features = Xp.shape[1]
        self._check_init(n_features)
        self._check_step(n_features)
        self.components_ = None
        self.continue_iteration = True
        self.num_iter = 0
        self._first_gram_ = None
        self._first_gradients_dev = None
        X_proj = Xp
        while self.continue_iteration:
            self._fit_estimator(X_proj, yd, **opt_kws)
            self._fit_transformer(X_proj)
            X_proj = self._project_rows(Xp)
            self.num_iter += 1
        self._last_fit(X_proj, yd, **opt_kws)

    def _fit_estimator(self, Xrows, yd, **opt_kws):
        """Clone and fit the estimator on the current (projected) rows (edrgp/base.py:115-140)."""
        self.estimator_ = clone(self.estimator)
        self.estimator_.fit(Xrows, yd, **opt_kws)
        if not hasattr(getattr(self.estimator_, 'estimator_', None), 'gradient_gram'):
            raise TypeError("estimator must provide device gradients (edrgp_b200.SparseGaussianProcessRegressor)")
        if self.num_iter == 0:
            self.first_estimator_ = clone(self.estimator_)
        return self

    def _gradient_gram(self, want_G):
        """(G on device or None, C reduced over ranks) of the current estimator on its training rows,
        mapped back through the preprocessor on the first iteration (edrgp/edr.py:233-238)."""
        G, C = self.estimator_.estimator_.gradient_gram(want_G=want_G, want_C=True, reduce=True)
        C = C.cpu().numpy()
        if self.preprocessor is not None and self.num_iter == 0:
            P = self._preprocessing_
            C = P.T.dot(C).dot(P)
            if G is not None:
                G = G @ torch.as_tensor(np.ascontiguousarray(P), device=G.device)
        return G, C

    def _fit_transformer(self, Xrows):
        """Fit the transformer on the gradients and select components (edrgp/base.py:468-517)."""
        check_is_fitted(self, 'estimator_')
        fused = hasattr(self.dr_transformer, 'fit_gram')
        want_G = (self.keep_gradients and self.num_iter == 0) or not fused
        G, C = self._gradient_gram(want_G)
        if self.num_iter == 0:
            self._first_gram_ = C
            self._first_gradients_dev = G if self.keep_gradients else None
        self.transformer_ = clone(self.dr_transformer)
        if fused:
            cnt = torch.tensor([float(self._n_local)], dtype=F64, device='cuda')
            dist.allreduce_sum_(cnt)
            self.transformer_.fit_gram(C, int(round(float(cnt[0]))))
        else:
            if dist.is_distributed():
                raise NotImplementedError("a host transformer needs all gradients in one process; "
                                          "use GramEighTransformer")
            self.transformer_.fit(G.cpu().numpy())
        self._check_transformer(self.transformer_)
        comps = deepcopy(self.transformer_.components_)
        n_components = self._select_n_components(C, comps)
        self.components_ = self._select_best_components(comps, n_components)
        return self

    def _select_n_components(self, C, components):
        dim = C.shape[0]
        if self.adaptive_step:
            _, var_ratio_ = subspace_variance_ratio_from_gram(C, components.T)
            n_components = int(np.sum(np.cumsum(var_ratio_) < self.step_, dtype=int)) + 1
            if n_components == dim:
                self.continue_iteration = False
        else:
            n_components = max(self.n_components_, dim - self.step_)
            if n_components == self.n_components_:
                self.continue_iteration = False
        return n_components

    def _select_best_components(self, components, n_components):
        self.components_ = components if self.components_ is None else np.dot(components, self.components_)
        _, var_ratio = subspace_variance_ratio_from_gram(self._first_gram_, self.components_.T)
        best_components = np.argsort(var_ratio)[-n_components:][::-1]
        return self.components_[best_components, :]

    def _project_rows(self, Xp):
        """In-loop ``transform`` of the preprocessed rows (edrgp/base.py:462, edrgp/edr.py:284-288)."""
        comps = self.components_
        if self.preprocessor is not None:
            comps = np.dot(comps, self._preprocessing_.T)
        if Xp.shape[0] == 0:
            return torch.empty(0, comps.shape[0], dtype=F64, device=Xp.device)
        V = torch.as_tensor(np.ascontiguousarray(comps, dtype=np.float64), device=Xp.device)
        return ops.project(Xp, V)

    def _last_fit(self, X_proj, yd, **opt_kws):
        """Fit the estimator on the projected rows and compute the subspace variance against the
        first gradients (edrgp/base.py:172-200)."""
        self._fit_estimator(X_proj, yd, **opt_kws)
        (self.subspace_variance_,
         self.subspace_variance_ratio_) = subspace_variance_ratio_from_gram(self._first_gram_, self.components_.T)
        self._components_fit_space = self.components_.copy()
        self._subspace_gradients_host = None
        return self

    # ------------------------------------------------------------------------------------------
    # gradients kept for refit / inspection (edrgp/base.py:161, 193-195)
    # ------------------------------------------------------------------------------------------
    @property
    def _first_gradients_(self):
        """Host copy of the first iteration's gradients (edrgp/base.py:161).  With several ranks every rank holds
        only ITS rows: that is what this returns, with a warning -- use ``first_gradients_rows`` for rows of the
        whole data set, or ``refit``, which gathers what it needs."""
        if getattr(self, '_first_gradients_dev', None) is None:
            raise AttributeError("first gradients were not kept (keep_gradients=False)")
        if dist.is_distributed():
            warnings.warn("_first_gradients_ holds this rank's %d rows only (the gradients are sharded over %d ranks); "
                          "first_gradients_rows(rows) gathers rows of the whole data set"
                          % (self._first_gradients_dev.shape[0], dist.world_size()), RuntimeWarning)
        return self._first_gradients_dev.cpu().numpy()

    def _row_offsets(self):
        """(first global row of this rank, global row count) of the row-sharded gradients."""
        sizes = torch.zeros(dist.world_size(), dtype=F64, device='cuda')
        sizes[dist.rank()] = float(self._n_local)
        dist.allreduce_sum_(sizes)
        sizes = sizes.cpu().numpy()
        return int(round(sizes[:dist.rank()].sum())), int(round(sizes.sum()))

    def first_gradients_rows(self, rows=None, max_rows=None, seed=0):
        """Rows of the first iteration's gradients by GLOBAL row index, gathered to every rank as a host array
        (k, d).  ``rows=None``: all rows, or -- when ``max_rows`` is given and smaller -- a seeded uniform subsample
        without replacement (sorted).  Returns ``(G_rows, rows)``.  Collective: every rank calls it alike."""
        if getattr(self, '_first_gradients_dev', None) is None:
            raise AttributeError("first gradients were not kept (keep_gradients=False)")
        G = self._first_gradients_dev
        lo, n = (0, G.shape[0]) if not dist.is_distributed() else self._row_offsets()
        if rows is None:
            rows = np.arange(n)
            if max_rows is not None and n > max_rows:
                rows = np.sort(np.random.RandomState(seed).choice(n, size=int(max_rows), replace=False))
        else:
            rows = np.asarray(rows)
            if rows.dtype == bool:
                rows = np.nonzero(rows)[0]
            rows = np.where(rows < 0, rows + n, rows)
        if not dist.is_distributed():
            idx = torch.as_tensor(rows, device=G.device)
            return G[idx].cpu().numpy(), rows
        out = torch.zeros(len(rows), G.shape[1], dtype=F64, device=G.device)
        local = rows - lo
        mine = np.nonzero((local >= 0) & (local < G.shape[0]))[0]
        if mine.size:
            out[torch.as_tensor(mine, device=G.device)] = G[torch.as_tensor(local[mine], device=G.device)]
        dist.allreduce_sum_(out)                      # every row is owned by exactly one rank
        return out.cpu().numpy(), rows

    @property
    def subspace_gradients_(self):
        """Gradients of the final estimator on its (projected) training rows, (n, k)."""
        check_is_fitted(self, 'estimator_')
        if self._subspace_gradients_host is None:
            G, _ = self.estimator_.estimator_.gradient_gram(want_G=True, want_C=False)
            self._subspace_gradients_host = G.cpu().numpy()
        return self._subspace_gradients_host

    @property
    def _recovered_gradients_(self):
        return np.dot(self.subspace_gradients_, self._components_fit_space)

    def refit(self, refit_transformer, rows=None, max_rows=1_000_000):
        """New components from the gradients kept at fit time (edrgp/edr.py:115-140,
        edrgp/base.py:202-239).  ``refit_transformer`` is any transformer with ``fit`` + ``components_``.

        * A transformer with ``fit_gram`` (``GramEighTransformer``) and ``rows=None`` is fitted from the Gram matrix of
          ALL gradients, which the fit has already reduced over rows and ranks: nothing is gathered.
        * A host transformer (e.g. ``SparsePCA``) needs rows: ``rows`` (global indices) if given, all rows if they are
          at most ``max_rows``, else a seeded uniform subsample of ``max_rows`` rows -- announced by a warning.  With
          several ranks the rows are gathered to every rank and every rank fits the same transformer (so give it a
          ``random_state``); the subspace variance is always taken against the rows the transformer saw."""
        check_is_fitted(self, 'components_')
        self.refit_transformer_ = clone(refit_transformer)
        if hasattr(self.refit_transformer_, 'fit_gram') and rows is None:
            cnt = torch.tensor([float(self._n_local)], dtype=F64, device='cuda')
            dist.allreduce_sum_(cnt)
            self.refit_transformer_.fit_gram(self._first_gram_, int(round(float(cnt[0]))))
            C = self._first_gram_
        else:
            grads, used = self.first_gradients_rows(rows, max_rows=max_rows if rows is None else None)
            if rows is None and dist.is_distributed():
                total = self._row_offsets()[1]
            else:
                total = self._first_gradients_dev.shape[0]
            if rows is None and len(used) < total:
                warnings.warn("refit: the host transformer is fitted on a seeded subsample of %d of %d gradient rows "
                              "(max_rows=%d)" % (len(used), total, max_rows), RuntimeWarning)
            self.refit_rows_ = used
            self.refit_transformer_.fit(grads)
            C = grads.T.dot(grads)
        self._check_transformer(self.refit_transformer_)
        comps = deepcopy(self.refit_transformer_.components_)
        comps = _l2_normalize(comps, axis=1)
        comps = self._remove_zero_components(comps)
        (self.refit_subspace_variance_,
         self.refit_subspace_variance_ratio_) = subspace_variance_ratio_from_gram(C, comps.T)
        self.refit_components_ = np.dot(comps, self._reverse_scaling_) if self.normalize else comps
        return self

    def _remove_zero_components(self, components):
        nonzero_indices = np.nonzero(np.linalg.norm(components, axis=1))[0]
        zero_components = sorted(set(range(components.shape[0])) - set(nonzero_indices))
        if zero_components:
            warnings.warn('Components with numbers {} will be droped because they '
                          'contains only zeros'.format(zero_components), RuntimeWarning)
        return np.delete(components, zero_components, axis=0)

    # ------------------------------------------------------------------------------------------
    # public surface (edrgp/edr.py:199-289, edrgp/base.py:302-319)
    # ------------------------------------------------------------------------------------------
    def get_estimator_gradients(self, X):
        """Gradients of the final estimator at new rows, in raw feature space (edrgp/edr.py:199-241)."""
        X = check_array(X)
        Xk = self._preprocessing_transform(X)
        check_is_fitted(self, 'estimator_')
        grad = self.estimator_.predict_gradient(Xk)
        return np.dot(grad, self.components_)

    @property
    def feature_importances_(self):
        check_is_fitted(self, 'components_')
        importances_ = self.components_
        if self.normalize is True:
            importances_ = np.dot(importances_, self._scaling_)
        return importances_

    def transform(self, X, refitted=False):
        """Project X on the EDR directions: a pure linear map, no centring (edrgp/edr.py:261-289).
        Host array in -> host array out; CUDA tensor in -> CUDA tensor out (``edrgp_project``)."""
        check_is_fitted(self, 'components_')
        if refitted:
            check_is_fitted(self, ['refit_transformer_', 'refit_components_'])
            comps = self.refit_components_
        else:
            comps = self.components_
        if isinstance(X, torch.Tensor):
            V = torch.as_tensor(np.ascontiguousarray(comps, dtype=np.float64), device=X.device)
            return ops.project(X.to(dtype=F64).contiguous(), V)
        X = check_array(X)
        return np.dot(X, comps.T)

    def inverse_transform(self, X):
        check_is_fitted(self, 'components_')
        X = check_array(X)
        return np.dot(X, np.linalg.pinv(self.components_).T)


EDR = EffectiveDimensionalityReduction


class BlockEDR(EffectiveDimensionalityReduction):
    """Block effective dimensionality reduction (``edrgp.base.BlockEDR``, edrgp/base.py:520-766): one transformer
    per block of features, fitted on the block's columns of the gradients, merged into block-diagonal components.

    The gradients of a block's columns have the Gram matrix ``C[cols, cols]``, so with a ``fit_gram`` transformer
    (``GramEighTransformer``) every block is a small eigenproblem on a diagonal block of the d x d Gram matrix the
    fused gradient kernel has already reduced over rows and ranks; other transformers are
    fitted on the gathered gradient rows.  Like the reference class it is one-shot (no ``step``) and works on the
    rows as given (no scaling, no preprocessor).

    Parameters: ``estimator``, ``transformer``, ``n_components`` (None, an int, or a list with one entry per
    block), ``blocks`` (list of lists of column indices; None = one block of all columns)."""

    def __init__(self, estimator=None, transformer=None, n_components=None, blocks=None, keep_gradients=True):
        self.estimator = estimator
        self.transformer = transformer
        self.n_components = n_components
        self.blocks = blocks
        self.keep_gradients = keep_gradients

    # the parent's fit machinery reads these (``transformer`` is a read-only alias there; here it is the parameter)
    transformer = None
    dr_transformer = property(lambda self: self.transformer)
    step = None
    normalize = False
    preprocessor = None

    def _check_init(self, n_features):
        if self.estimator is None:
            raise ValueError("Estimator should be speciified")
        if self.transformer is None:
            raise ValueError("transformer should be specified")
        self.n_components_ = n_features if self.n_components is None else self.n_components

    def _make_blocks(self, n_features):                       # edrgp/base.py:735-766
        if self.blocks is None:
            if isinstance(self.n_components_, (int, np.integer)):
                self.blocks_ = [{'columns': np.arange(n_features), 'n_components': int(self.n_components_)}]
            else:
                raise ValueError("blocks should be specified if n_components is list")
        elif isinstance(self.blocks, list):
            if isinstance(self.n_components_, list):
                self.blocks_ = [{'columns': np.asarray(b), 'n_components': k}
                                for b, k in zip(self.blocks, self.n_components_)]
            else:
                # (the reference takes max(n_components, len(block)): an int keeps every direction of every block)
                self.blocks_ = [{'columns': np.asarray(b), 'n_components': max(int(self.n_components_), len(b))}
                                for b in self.blocks]
        else:
            raise ValueError("blocks should be None or a list of lists of column indices")
        return self

    def _fit_block(self, transformer, C, block, grads=None):
        transformer.set_params(n_components=block['n_components'])
        cols = block['columns']
        if hasattr(transformer, 'fit_gram'):
            transformer.fit_gram(C[np.ix_(cols, cols)], self._n_total)
        else:
            transformer.fit(grads[:, cols])
        self._check_transformer(transformer)
        return transformer.components_.T

    def _merge_components(self, components, n_features):      # edrgp/base.py:654-680
        eff = sum(c.shape[1] for c in components)
        out = np.zeros((n_features, eff))
        start = 0
        for blk, comp in zip(self.blocks_, components):
            stop = start + comp.shape[1]
            out[blk['columns'], start:stop] = comp
            blk['components'] = np.arange(start, stop)
            start = stop
        return out.T

    def _fit_iterations(self, Xp, yd, **opt_kws):
        n_features = Xp.shape[1]
        self._check_init(n_features)
        self._make_blocks(n_features)
        self.components_ = None
        self.num_iter = 0
        self._first_gram_ = None
        self._first_gradients_dev = None
        cnt = torch.tensor([float(self._n_local)], dtype=F64, device='cuda')
        dist.allreduce_sum_(cnt)
        self._n_total = int(round(float(cnt[0])))
        self._fit_estimator(Xp, yd, **opt_kws)
        fused = hasattr(self.transformer, 'fit_gram')
        G, C = self._gradient_gram(self.keep_gradients or not fused)
        self._first_gram_ = C
        self._first_gradients_dev = G if self.keep_gradients else None
        grads = None
        if not fused:
            self._first_gradients_dev = G
            grads = self.first_gradients_rows(None)[0]
            if not self.keep_gradients:
                self._first_gradients_dev = None
        self.transformer_ = clone(self.transformer)
        comps = [self._fit_block(clone(self.transformer), C, blk, grads) for blk in self.blocks_]
        self.components_ = self._merge_components(comps, n_features)
        X_proj = self._project_rows(Xp)
        self.num_iter += 1
        self._last_fit(X_proj, yd, **opt_kws)

    def refit(self, refit_transformer, rows=None, params=None, max_rows=1_000_000):
        """Per-block refit on the kept gradients (edrgp/base.py:682-733); ``params``: one dict of extra transformer
        parameters per block."""
        check_is_fitted(self, 'components_')
        self._make_blocks(self._first_gram_.shape[0])
        self.refit_transformer_ = clone(refit_transformer)
        fused = hasattr(refit_transformer, 'fit_gram') and rows is None
        if fused:
            grads, C = None, self._first_gram_
        else:
            grads, self.refit_rows_ = self.first_gradients_rows(rows, max_rows=max_rows if rows is None else None)
            C = grads.T.dot(grads)
        comps = []
        for i, blk in enumerate(self.blocks_):
            tr = clone(refit_transformer)
            if params is not None and params[i] is not None:
                tr.set_params(**params[i])
                blk = dict(blk, n_components=params[i].get('n_components', blk['n_components']))
            comps.append(self._fit_block(tr, C, blk, grads))
        merged = _l2_normalize(self._merge_components(comps, C.shape[0]))
        self.refit_components_ = self._remove_zero_components(merged)
        (self.refit_subspace_variance_,
         self.refit_subspace_variance_ratio_) = subspace_variance_ratio_from_gram(C, self.refit_components_.T)
        return self
