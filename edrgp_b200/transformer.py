"""``GramEighTransformer``: drop-in for edr-gp's ``SVDTransformer`` (``edrgp/utils.py:81-175``,
"PCA without centering and scaling") that never forms the n x n ``U`` of a full SVD.

The right singular vectors of G are the eigenvectors of C = G^T G and S^2 its eigenvalues, so
``fit(G)`` reduces G to the d x d Gram matrix on the FP64 tensor pipe (``edrgp_syrk``) and
eigendecomposes it with the Jacobi kernel (``edrgp_eigh``).  ``fit_gram(C)`` takes a Gram matrix
that is already on hand -- the fused gradient kernel produces it without G ever leaving the chip.
"""
import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array

from . import dist, ops

F64 = torch.float64


class GramEighTransformer(BaseEstimator, TransformerMixin):
    """Linear dimensionality reduction without centring.

    Parameters
    ----------
    n_components : int, float or None
        As ``SVDTransformer``: None keeps all; an int keeps that many; a float in (0, 1) keeps the
        smallest number of components whose cumulative variance ratio reaches it.

    Attributes
    ----------
    components_ : (n_components, n_features)   rows = directions, descending variance
    subspace_variance_ : (n_components,)        S^2 of the gradients (eigenvalues of G^T G)
    subspace_variance_ratio_ : (n_components,)  S^2 / sum(S^2)
    gram_ : (n_features, n_features)            the Gram matrix that was decomposed
    """

    def __init__(self, n_components=None):
        self.n_components = n_components

    def fit(self, X, y=None):
        """Fit on gradients X (n, d): host array or CUDA tensor (this rank's rows)."""
        if isinstance(X, torch.Tensor):
            Xd = X.to(dtype=F64).contiguous()
            if not Xd.is_cuda:
                Xd = Xd.cuda()
        else:
            X = check_array(X, dtype=np.float64)
            Xd = torch.as_tensor(np.ascontiguousarray(X), device='cuda')
        n, d = Xd.shape
        C = ops.syrk(ops.pad_even(Xd))[:d, :d].contiguous()
        cnt = torch.tensor([float(n)], dtype=F64, device=C.device)
        dist.allreduce_sum_(C, cnt)
        return self.fit_gram(C, int(round(float(cnt[0]))))

    def fit_gram(self, C, n_samples=None):
        """Fit from C = G^T G (d, d), already summed over all rows (and ranks)."""
        fixed = getattr(C, '_edrgp_fixed', None)
        fixed = fixed() if fixed is not None else None
        if fixed is not None and tuple(C.shape) == (fixed.d, fixed.d):
            # C came from a composite sweep (SparseGPRegression.gradient_gram): the eigensolver runs inside the
            # sweep's result block and ONE read-back brings eigenvalues, components, C and the sweep's
            # deferred-check words to the host
            d = fixed.d
            host = fixed.eigh(C)
        else:
            if not isinstance(C, torch.Tensor):
                C = torch.as_tensor(np.ascontiguousarray(C, dtype=np.float64), device='cuda')
            d = C.shape[0]
            evals, comps, status = ops.eigh(C, return_status=True)
            # one read-back for eigenvalues, eigenvectors, the Gram matrix itself and the solver's status
            host = torch.cat([evals, comps.reshape(-1), C.reshape(-1), status.to(F64)]).cpu().numpy()
            ops.check_eigh_status(host[-2], host[-1])
        S2 = np.clip(host[:d], 0.0, np.inf)
        comps = host[d:d + d * d].reshape(d, d)
        gram = host[d + d * d:d + 2 * d * d].reshape(d, d)
        total = S2.sum()
        ratio = S2 / total if total > 0 else np.zeros_like(S2)
        nc = self.n_components
        k = d
        if nc is None:
            k = d
        elif isinstance(nc, (int, np.integer)) and 0 < nc <= d:
            k = int(nc)
        elif isinstance(nc, float) and 0 < nc < 1:
            k = int(np.sum(np.cumsum(ratio) < nc, dtype=int)) + 1
        if n_samples is not None:
            k = min(int(n_samples), k)
        self.gram_ = gram
        self.components_ = comps[:k, :]
        self.subspace_variance_ = S2[:k]
        self.subspace_variance_ratio_ = ratio[:k]
        return self

    def transform(self, X):
        """Project X (n, d) on the components: host in, host out; CUDA tensor in, CUDA tensor out."""
        if isinstance(X, torch.Tensor):
            V = torch.as_tensor(np.ascontiguousarray(self.components_), device=X.device)
            return ops.project(X.contiguous(), V)
        return np.asarray(X).dot(self.components_.T)


class DevicePCA(BaseEstimator, TransformerMixin):
    """PCA preprocessor on the device: drop-in for ``sklearn.decomposition.PCA``
    as edr-gp uses it in front of the estimator (``preprocessor=PCA(n_components=...)``,
    ``edrgp/edr.py:169-174``; examples/chain_PCA-EDRGP.ipynb).

    The column means and the centred Gram matrix X_c^T X_c are reduced on the device (column-moment
    and DMMA symmetric-reduction kernels, summed over ranks), the d x d eigenproblem runs on the Jacobi
    kernel and the rows are projected by ``edrgp_project``: X never visits the host.  Attributes follow
    sklearn: ``components_`` (k, d) with the largest-magnitude entry of each row positive, ``mean_``,
    ``explained_variance_`` (ddof = 1), ``explained_variance_ratio_``, ``singular_values_``.
    """

    def __init__(self, n_components=None):
        self.n_components = n_components

    def _fit_device(self, Xd):
        n_local, d = Xd.shape
        cnt = torch.tensor([float(n_local)], dtype=F64, device=Xd.device)
        s1 = ops.col_moments(Xd)[0].clone() if n_local else torch.zeros(d, dtype=F64, device=Xd.device)
        dist.allreduce_sum_(s1, cnt)
        n = float(cnt[0])
        mean = s1 / n
        ones = torch.ones(d, dtype=F64, device=Xd.device)
        Xc = ops.standardize(Xd, mean, ones) if n_local else Xd
        if n_local:
            C = ops.syrk(ops.pad_even(Xc))[:d, :d].contiguous()
        else:
            C = torch.zeros(d, d, dtype=F64, device=Xd.device)
        dist.allreduce_sum_(C)
        evals, comps, status = ops.eigh(C, return_status=True)
        host = torch.cat([evals, comps.reshape(-1), status.to(F64)]).cpu().numpy()
        ops.check_eigh_status(host[-2], host[-1])
        lam = np.clip(host[:d], 0.0, np.inf)
        comps = host[d:d + d * d].reshape(d, d)
        nc = self.n_components
        if nc is None:
            k = min(d, int(n))
        elif isinstance(nc, (int, np.integer)) and 0 < nc <= d:
            k = int(nc)
        elif isinstance(nc, float) and 0 < nc < 1:
            ratio = lam / lam.sum()
            k = int(np.searchsorted(np.cumsum(ratio), nc, side='right')) + 1
            k = min(k, d)
        else:
            raise ValueError("n_components=%r is not supported" % (nc,))
        self.mean_ = mean.cpu().numpy()
        self.n_samples_ = int(round(n))
        self.n_features_in_ = d
        self.components_ = comps[:k]
        self.explained_variance_ = lam[:k] / max(n - 1.0, 1.0)
        self.explained_variance_ratio_ = lam[:k] / lam.sum() if lam.sum() > 0 else np.zeros(k)
        self.singular_values_ = np.sqrt(lam[:k])
        self.n_components_ = k
        return Xc

    def fit_transform_device(self, Xd):
        """Fit on this rank's device rows and return their projection (device tensor, (n_local, k))."""
        Xc = self._fit_device(Xd.contiguous())
        if Xc.shape[0] == 0:
            return torch.empty(0, self.n_components_, dtype=F64, device=Xd.device)
        return ops.project(Xc, torch.as_tensor(np.ascontiguousarray(self.components_), device=Xd.device))

    def fit(self, X, y=None):
        self._fit_device(_rows_to_device(X))
        return self

    def fit_transform(self, X, y=None):
        out = self.fit_transform_device(_rows_to_device(X))
        return out if isinstance(X, torch.Tensor) else out.cpu().numpy()

    def transform(self, X):
        if isinstance(X, torch.Tensor):
            mean = torch.as_tensor(self.mean_, device=X.device)
            Xc = ops.standardize(X.to(dtype=F64).contiguous(), mean, torch.ones_like(mean))
            return ops.project(Xc, torch.as_tensor(np.ascontiguousarray(self.components_), device=X.device))
        return (np.asarray(X) - self.mean_).dot(self.components_.T)


def _rows_to_device(X):
    if isinstance(X, torch.Tensor):
        Xd = X.to(dtype=F64)
        return (Xd if Xd.is_cuda else Xd.cuda()).contiguous()
    X = check_array(X, dtype=np.float64)
    return torch.as_tensor(np.ascontiguousarray(X), device='cuda')
