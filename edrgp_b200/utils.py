"""Host-side helpers on d x d matrices: the Gram-matrix forms of edr-gp's ``subspace_variance_ratio``
(``edrgp/utils.py:27-55``) and ``discrepancy`` (``edrgp/utils.py:58-78``).  Everything here is
O(d^2 k) on the host; the n-scale work (the Gram matrix itself) is done on the device."""
import numpy as np


def subspace_variance_ratio_from_gram(C, V):
    """``subspace_variance_ratio(G, V)`` from C = G^T G: per-column ||G v_k|| (orthonormal V) or one
    Frobenius norm ||G Q|| with Q = qr(V) otherwise; ratio = (norm / ||G||_F)^2."""
    C = np.asarray(C, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    if np.allclose(np.dot(V.T, V), np.eye(V.shape[1])):
        var = np.sqrt(np.clip(np.einsum('ik,ij,jk->k', V, C, V), 0, np.inf))
    else:
        Q = np.linalg.qr(V)[0]
        var = np.sqrt(max(float(np.trace(Q.T.dot(C).dot(Q))), 0.0))
    ratio = var ** 2 / np.trace(C)
    return var, ratio


def subspace_variance_ratio(X, V):
    """The row form the reference exposes (``edrgp/utils.py:27-55``): X (n, d) gradients -- a host array or a CUDA
    tensor holding this rank's rows -- and V (d, k).  The n-scale part, C = X^T X, is reduced on the device
    (``edrgp_syrk``, summed over ranks); the rest is ``subspace_variance_ratio_from_gram``."""
    from . import dist, ops
    from .transformer import _rows_to_device
    Xd = _rows_to_device(X)
    d = Xd.shape[1]
    C = ops.syrk(ops.pad_even(Xd))[:d, :d].contiguous()
    dist.allreduce_sum_(C)
    return subspace_variance_ratio_from_gram(C.cpu().numpy(), V)


def ort_space(A, tol=1e-10):
    """Orthonormal basis (n_features, n_features - rank) of the complement of span(A), A (n_features, k):
    the left singular vectors behind the singular values above ``tol`` (``edrgp/utils.py:8-24``)."""
    A = np.asarray(A, dtype=np.float64)
    U, s, _ = np.linalg.svd(A, full_matrices=True)
    return U[:, int(np.count_nonzero(np.abs(s) > tol)):]


def discrepancy(B, V):
    """||B B^T (I - V V^T)||_F / d_true for a true projector basis B (n_features, d_true) and an
    estimated one V (n_features, k), as edrgp/utils.py:58-78."""
    B = np.asarray(B, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    return float(np.linalg.norm(B.dot(B.T).dot(np.eye(B.shape[0]) - V.dot(V.T)))) / B.shape[1]


def principal_angle(A, B):
    """Largest principal angle (radians) between the row spaces of A and B (k, d)."""
    Qa = np.linalg.qr(np.asarray(A).T)[0]
    Qb = np.linalg.qr(np.asarray(B).T)[0]
    R = Qb - Qa.dot(Qa.T.dot(Qb))
    s = np.linalg.svd(R, compute_uv=False)
    return float(np.arcsin(min(1.0, s.max())))
