"""Fixed-hyper-parameter units of the EDR hot path on the CPU.  TEST ORACLE / CPU BASELINE.

Two flavours of every unit (units K1-K6):

* ``*_faithful``: calls ``oracle.gpy_restatement`` exactly as GPy would be called from
  ``edrgp/gp_model/base.py:65-69,222`` and ``edrgp/utils.py:140`` (n x m temporaries, per-dimension
  gradient loop, full SVD).  O(n m) / O(n^2) memory: small n only.
* ``*_chunked``: the same arithmetic regrouped so it runs at the BASELINE shapes -- 16 384-row
  chunks, GEMM-form gradient ``G = ((Kfu o a^T) Z - rowsum(Kfu o a^T) o X) / l^2``, ``G^T G``
  accumulated per chunk, ``eigh`` at the end.  This is what
  ``bench.py`` times as ``cpu_baseline`` / ``--impl reference`` (kind "port").

Nothing under ``edrgp_b200/`` imports this module.
"""
import numpy as np

from . import gpy_restatement as gpy

CHUNK = 16384


# ------------------------------------------------------------------------------------------------
# K1: cross-covariance
# ------------------------------------------------------------------------------------------------
def kuf_faithful(X, Z, ell, sf2):
    """Kfu (n, m) through the restated ``RBF.K`` (ARD)."""
    kern = gpy.RBF(X.shape[1], variance=sf2, lengthscale=np.asarray(ell, float), ARD=True)
    return kern.K(X, Z)


def _kfu_chunk(Xc, Zs, zn, ell, sf2):
    Xs = Xc / ell
    r2 = -2. * Xs.dot(Zs.T) + (np.sum(np.square(Xs), 1)[:, None] + zn[None, :])
    np.clip(r2, 0, np.inf, out=r2)
    return r2, Xs


# ------------------------------------------------------------------------------------------------
# K2: inducing statistics  P = Kuf Kfu, b = Kuf y, yy = y^T y
# ------------------------------------------------------------------------------------------------
def inducing_stats_chunked(X, y, Z, ell, sf2, chunk=CHUNK):
    ell = np.asarray(ell, float)
    m = Z.shape[0]
    Zs = Z / ell
    zn = np.sum(np.square(Zs), 1)
    P = np.zeros((m, m))
    b = np.zeros(m)
    yy = 0.0
    for s in range(0, X.shape[0], chunk):
        r2, _ = _kfu_chunk(X[s:s + chunk], Zs, zn, ell, sf2)
        K = sf2 * np.exp(-0.5 * r2)
        P += K.T.dot(K)
        yc = y[s:s + chunk]
        b += K.T.dot(yc)
        yy += float(yc.dot(yc))
    return P, b, yy


# ------------------------------------------------------------------------------------------------
# K3: the m x m solve chain of VarDTC, driven by the statistics instead of the n x m matrix
# ------------------------------------------------------------------------------------------------
def kuu(Z, ell, sf2, jitter=gpy.CONST_JITTER):
    kern = gpy.RBF(Z.shape[1], variance=sf2, lengthscale=np.asarray(ell, float), ARD=True)
    Kmm = kern.K(Z).copy()
    Kmm[np.diag_indices(Z.shape[0])] += jitter
    return Kmm


def solve_from_stats(Kmm, P, b, yy, n, sf2, noise_variance):
    """alpha (woodbury vector), woodbury_inv and the VFE bound from {Kuu, P, b, yy, n}.

    Same Cholesky chain as ``gpy.vardtc_inference`` with ``A = beta Lm^-1 P Lm^-T`` replacing
    ``tdot(Lm^-1 Kuf sqrt(beta))`` and ``Lm^-1 b`` replacing ``Lm^-1 Kuf y`` -- algebraically
    identical; this is the form an n-sharded implementation has to use.
    """
    m = Kmm.shape[0]
    beta = 1. / max(float(noise_variance), gpy.CONST_JITTER)
    Lm = gpy.jitchol(Kmm)
    A = beta * gpy.backsub_both_sides(Lm, P, transpose='right')       # Lm^-1 P Lm^-T
    A = 0.5 * (A + A.T)
    B = np.eye(m) + A
    LB = gpy.jitchol(B)
    c = gpy.dtrtrs(LB, gpy.dtrtrs(Lm, b[:, None], lower=1), lower=1) * beta   # LB^-1 Lm^-1 Kuf (beta y)
    tmp = gpy.dtrtrs(LB, c, lower=1, trans=1)
    alpha = gpy.dtrtrs(Lm, tmp, lower=1, trans=1)[:, 0]
    data_fit = float(np.sum(np.square(c)))
    bound = (-0.5 * n * (np.log(2. * np.pi) - np.log(beta)) - 0.5 * beta * yy
             - 0.5 * (beta * n * sf2 - np.trace(A))
             - np.sum(np.log(np.diag(LB))) + 0.5 * data_fit)
    Bi = -gpy.dpotri(LB)
    Bi[np.diag_indices(m)] += 1
    woodbury_inv = gpy.backsub_both_sides(Lm, Bi)
    return {'alpha': alpha, 'woodbury_inv': woodbury_inv, 'bound': float(bound),
            'Lm': Lm, 'LB': LB, 'A': A}


# ------------------------------------------------------------------------------------------------
# K4: posterior-mean gradients
# ------------------------------------------------------------------------------------------------
def gradients_faithful(X, Z, ell, sf2, alpha, scale=1.0):
    """``Stationary.gradients_X(alpha^T, X, Z)``: per-dimension loop over n x m temporaries."""
    kern = gpy.RBF(X.shape[1], variance=sf2, lengthscale=np.asarray(ell, float), ARD=True)
    return kern.gradients_X(np.asarray(alpha)[None, :], X, Z) * scale


def gradients_chunk(Xc, Z, Zs, zn, ell, sf2, alpha, scale=1.0):
    r2, _ = _kfu_chunk(Xc, Zs, zn, ell, sf2)
    W = sf2 * np.exp(-0.5 * r2)
    W *= alpha[None, :]
    W[r2 == 0.] = 0.                      # GPy's _inv_dist zeroes coincident pairs
    G = W.dot(Z)
    G -= W.sum(1)[:, None] * Xc
    G /= ell ** 2
    if scale != 1.0:
        G *= scale
    return G


def gradients_chunked(X, Z, ell, sf2, alpha, scale=1.0, chunk=CHUNK):
    ell = np.asarray(ell, float)
    Zs = Z / ell
    zn = np.sum(np.square(Zs), 1)
    G = np.empty_like(X)
    for s in range(0, X.shape[0], chunk):
        G[s:s + chunk] = gradients_chunk(X[s:s + chunk], Z, Zs, zn, ell, sf2, alpha, scale)
    return G


# ------------------------------------------------------------------------------------------------
# K5 + K6: gradient outer product and its eigendecomposition
# ------------------------------------------------------------------------------------------------
def grad_gram_chunked(X, Z, ell, sf2, alpha, scale=1.0, chunk=CHUNK):
    """C = G^T G (d, d) without keeping G: the 'Kuf + gradient + G^T G' pipeline on the CPU."""
    ell = np.asarray(ell, float)
    d = X.shape[1]
    Zs = Z / ell
    zn = np.sum(np.square(Zs), 1)
    C = np.zeros((d, d))
    for s in range(0, X.shape[0], chunk):
        G = gradients_chunk(X[s:s + chunk], Z, Zs, zn, ell, sf2, alpha, scale)
        C += G.T.dot(G)
    return C


def edr_from_gram(C, n_components=None):
    """``SVDTransformer.fit`` (edrgp/utils.py:123-157) restated on C = G^T G: right singular
    vectors of G are eigenvectors of C, S^2 its eigenvalues.  Rows sorted by descending variance;
    signs are arbitrary (as they are for LAPACK's SVD)."""
    C = 0.5 * (C + C.T)
    lam, V = np.linalg.eigh(C)
    order = np.argsort(lam)[::-1]
    lam = np.clip(lam[order], 0, np.inf)
    comps = V[:, order].T
    k = C.shape[0] if n_components is None else n_components
    ratio = lam / np.sum(lam)
    return comps[:k], lam[:k], ratio[:k]


def svd_faithful(G, n_components=None):
    """``SVDTransformer.fit`` with the n x n U avoided (economy SVD; same Vh and S)."""
    _, S, Vh = np.linalg.svd(G, full_matrices=False)
    k = G.shape[1] if n_components is None else n_components
    k = min(G.shape[0], k)
    return Vh[:k], (S ** 2)[:k], (S ** 2 / np.sum(S ** 2))[:k]


def principal_angle(A, B):
    """Largest principal angle (radians) between the row spaces of A and B (k, d)."""
    Qa = np.linalg.qr(A.T)[0]
    Qb = np.linalg.qr(B.T)[0]
    # sin-based formula, accurate for small angles
    R = Qb - Qa.dot(Qa.T.dot(Qb))
    s = np.linalg.svd(R, compute_uv=False)
    return float(np.arcsin(min(1.0, s.max())))


# ------------------------------------------------------------------------------------------------
# Synthetic workloads of SURVEY.md section 8(d) -- shared by tests and bench so both arms see the
# same inputs.  Vectorised (edrgp/datasets.py:39-57 uses Python list comprehensions).
# ------------------------------------------------------------------------------------------------
def make_workload(n, d, m, seed=0, k_true=3):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n, d))
    X -= X.mean(0)
    X /= X.std(0)
    B = np.linalg.qr(rng.standard_normal((d, k_true)))[0]
    y = np.tanh(X.dot(B)).sum(1) + 0.05 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    ell = np.sqrt(d) * (1. + 0.5 * np.random.RandomState(seed + 1).uniform(size=d))
    Z = X[rng.permutation(n)[:m]].copy()
    return {'X': X, 'y': y, 'Z': Z, 'ell': ell, 'sf2': 1.0, 'noise': 0.1, 'B': B}
