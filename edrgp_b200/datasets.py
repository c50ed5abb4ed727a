"""Synthetic workloads of the reference's examples and tests (``edrgp/datasets.py``), so that code written against
``edrgp.datasets`` keeps running after the import is switched.  Host-side NumPy on purpose: these are the INPUTS of
the path, and they draw from NumPy's global legacy generator in the reference's order, so ``np.random.seed(s)``
followed by the same calls gives the same arrays as the reference, bit for bit (``tests/test_support_surface.py``).
"""
import math

import numpy as np

__all__ = ['get_gaussian_inputs', 'get_tanh_targets', 'get_beta_inputs', 'get_edr_target', 'get_branin_targets']


def get_gaussian_inputs(sample_size, eig_values, eig_vectors=None, mean=None):
    """``sample_size`` draws of N(mean, Q diag(eig_values) Q^T) (``edrgp/datasets.py:7-22``); Q = ``eig_vectors``
    (used as given, not normalised) or a random rotation from ``scipy.stats.special_ortho_group``."""
    lam = np.asarray(eig_values, dtype=np.float64)
    k = lam.size
    if eig_vectors is None:
        from scipy.stats import special_ortho_group
        Q = special_ortho_group.rvs(k)
    else:
        Q = np.asarray(eig_vectors)
        if Q.ndim != 2 or not np.all(np.isfinite(Q)):
            raise ValueError("eig_vectors must be a finite 2-D array")
        if Q.shape != (k, k):
            raise ValueError('eig_vectors shape must be ({0},{0})'.format(k))
    cov = Q.dot(np.diag(lam)).dot(Q.T)
    return np.random.multivariate_normal(np.zeros(k) if mean is None else mean, cov, sample_size)


def get_tanh_targets(X, coefs, bias=0, noise_std=0.05):
    """tanh(X . coefs + bias) + noise_std * N(0, 1) (``edrgp/datasets.py:25-31``)."""
    if X.shape[1] != len(coefs):
        raise ValueError('Dimensionality of input ({}) and coefs ({}) are mismatched'.format(X.shape[1], len(coefs)))
    y = np.tanh(np.dot(X, coefs) + bias)
    y += noise_std * np.random.randn(X.shape[0])
    return y


def get_beta_inputs(sample_size, ndim, tau=1):
    """Entries from 2 Beta(1, tau) - 1 (``edrgp/datasets.py:34-36``)."""
    return 2 * np.random.beta(1, tau, size=(sample_size, ndim)) - 1


def get_edr_target(X, sigma=None):
    """The reference's test functions of 1, 2 or 3 effective coordinates (``edrgp/datasets.py:39-57``):
    u sin(sqrt(5) u);  (u1^3 + u2)(u1 - u2^3);  the same + u3; plus sigma * N(0, 1) when sigma is given."""
    U = np.asarray(X, dtype=np.float64)
    k = U.shape[1]
    if k == 1:
        g = U[:, 0] * np.sin(np.sqrt(5) * U[:, 0])
    elif k in (2, 3):
        u1, u2 = U[:, 0], U[:, 1]
        g = (u1 ** 3 + u2) * (u1 - u2 ** 3)
        if k == 3:
            g = g + U[:, 2]
    else:
        raise ValueError("get_edr_target is defined for 1, 2 or 3 effective coordinates (got %d)" % k)
    g = np.array(g, dtype=np.float64).ravel()
    if sigma is not None:
        g += sigma * np.random.randn(g.size)
    return g


def get_branin_targets(X, noise_std=None):
    """Branin function of the unit square rescaled to [-5, 10] x [0, 15] (``edrgp/datasets.py:60-93``)."""
    a, b, c = 1, 5.1 / (4 * math.pi ** 2), 5 / math.pi
    r, s, t = 6, 10, 1 / (8 * math.pi)
    x0 = 15 * X[:, 0] - 5
    x1 = 15 * X[:, 1]
    y = a * (x1 - b * x0 ** 2 + c * x0 - r) ** 2 + s * (1 - t) * np.cos(x0) + s
    if noise_std is not None:
        y += noise_std * np.random.randn(X.shape[0])
    return y
