"""Pins the CPU oracle: mathematical known-answer checks (no GPy needed) and the reference's own
tests (edrgp/tests/test_edr.py) re-run through the UNMODIFIED reference L3 with the oracle
estimator plugged in.  SURVEY.md section 4(ii)/(iii), section 8(c)."""
import numpy as np
import pytest

from oracle import gpy_restatement as gpy
from oracle import pipeline as op
from oracle.estimator import GaussianProcessRegressor, SparseGaussianProcessRegressor


def _small(n=400, d=6, m=25, seed=0):
    w = op.make_workload(n, d, m, seed=seed, k_true=2)
    return w


def _fit_fixed(w):
    kern = gpy.RBF(w['X'].shape[1], w['sf2'], w['ell'], ARD=True)
    post, ll, gd = gpy.vardtc_inference(kern, w['X'], w['Z'], w['noise'], w['y'][:, None])
    return kern, post, ll, gd


def test_alpha_matches_direct_solve():
    w = _small()
    kern, post, _, _ = _fit_fixed(w)
    Kuf = kern.K(w['Z'], w['X'])
    beta = 1 / w['noise']
    Kuu = kern.K(w['Z']) + 1e-8 * np.eye(w['Z'].shape[0])
    direct = np.linalg.solve(Kuu + beta * Kuf.dot(Kuf.T), beta * Kuf.dot(w['y']))
    assert np.allclose(post.woodbury_vector[:, 0], direct, rtol=1e-7, atol=1e-9)


def test_bound_matches_direct_formula():
    w = _small(n=300, m=20)
    kern, post, ll, _ = _fit_fixed(w)
    X, Z, y = w['X'], w['Z'], w['y']
    Kuu = kern.K(Z) + 1e-8 * np.eye(Z.shape[0])
    Kuf = kern.K(Z, X)
    Qff = Kuf.T.dot(np.linalg.solve(Kuu, Kuf))
    S = Qff + w['noise'] * np.eye(X.shape[0])
    sign, logdet = np.linalg.slogdet(S)
    direct = (-0.5 * X.shape[0] * np.log(2 * np.pi) - 0.5 * logdet - 0.5 * y.dot(np.linalg.solve(S, y))
              - 0.5 / w['noise'] * (X.shape[0] * w['sf2'] - np.trace(Qff)))
    assert ll.shape == (1, 1)
    assert abs(ll[0, 0] - direct) < 1e-8 * abs(direct)


def test_stats_form_equals_vardtc():
    w = _small()
    kern, post, ll, _ = _fit_fixed(w)
    P, b, yy = op.inducing_stats_chunked(w['X'], w['y'], w['Z'], w['ell'], w['sf2'], chunk=97)
    Kmm = op.kuu(w['Z'], w['ell'], w['sf2'])
    out = op.solve_from_stats(Kmm, P, b, yy, w['X'].shape[0], w['sf2'], w['noise'])
    assert np.allclose(out['alpha'], post.woodbury_vector[:, 0], rtol=1e-8, atol=1e-10)
    assert abs(out['bound'] - ll[0, 0]) < 1e-10 * abs(ll[0, 0])
    assert np.allclose(out['woodbury_inv'], post.woodbury_inv, rtol=1e-7, atol=1e-9)


def test_gradient_loop_equals_gemm_form_and_fd():
    w = _small()
    kern, post, _, _ = _fit_fixed(w)
    alpha = post.woodbury_vector[:, 0]
    Gf = op.gradients_faithful(w['X'], w['Z'], w['ell'], w['sf2'], alpha)
    Gc = op.gradients_chunked(w['X'], w['Z'], w['ell'], w['sf2'], alpha, chunk=64)
    assert np.max(np.abs(Gf - Gc)) <= 1e-12 * np.max(np.abs(Gf))
    # finite differences of mu(x) = K(x, Z) alpha
    eps = 1e-6
    i = 7
    for q in range(w['X'].shape[1]):
        xp = w['X'][i:i + 1].copy(); xm = xp.copy()
        xp[0, q] += eps; xm[0, q] -= eps
        fd = (kern.K(xp, w['Z']).dot(alpha) - kern.K(xm, w['Z']).dot(alpha))[0] / (2 * eps)
        assert abs(fd - Gf[i, q]) < 1e-7 * max(1.0, abs(Gf[i, q]))


def test_inducing_equals_data_is_dense_gp():
    rng = np.random.RandomState(3)
    X = rng.standard_normal((40, 2)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(40)
    kern = gpy.RBF(2, 1.3, np.array([0.9, 1.4]), ARD=True)
    post, _, _ = gpy.vardtc_inference(kern, X, X.copy(), 0.05, y[:, None])
    mu_sparse = kern.K(X, X).dot(post.woodbury_vector[:, 0])
    mu_dense = kern.K(X).dot(np.linalg.solve(kern.K(X) + 0.05 * np.eye(40), y))
    assert np.allclose(mu_sparse, mu_dense, atol=1e-5)


def test_oracle_matches_sklearn_gp_when_inducing_equals_data():
    """Independent third-party pin: with Z = X the variational sparse GP IS the exact GP, so the oracle's
    posterior mean, VFE bound, posterior-mean gradient and predictive variance must agree with
    scikit-learn's GaussianProcessRegressor (an implementation that shares no code with GPy or with this
    repository) on the same ARD-RBF kernel and noise."""
    from sklearn.gaussian_process import GaussianProcessRegressor as SkGP
    from sklearn.gaussian_process.kernels import RBF as SkRBF, ConstantKernel
    rng = np.random.RandomState(11)
    n, d = 60, 3
    X = rng.standard_normal((n, d)); y = np.tanh(X[:, 0] - 0.5 * X[:, 2]) + 0.05 * rng.standard_normal(n)
    sf2, ell, noise = 1.7, np.array([0.8, 1.9, 1.2]), 0.03
    kern = gpy.RBF(d, sf2, ell, ARD=True)
    post, ll, _ = gpy.vardtc_inference(kern, X, X.copy(), noise, y[:, None])
    sk = SkGP(kernel=ConstantKernel(sf2, 'fixed') * SkRBF(ell, 'fixed'), alpha=noise, optimizer=None).fit(X, y)
    Xs = rng.standard_normal((25, d))
    alpha = post.woodbury_vector[:, 0]
    mu = kern.K(Xs, X).dot(alpha)
    mu_sk, sd_sk = sk.predict(Xs, return_std=True)
    # GPy adds 1e-8 jitter to Kuu: agreement to ~1e-6, far above what a wrong formula would leave
    assert np.allclose(mu, mu_sk, atol=2e-6)
    assert abs(float(np.asarray(ll).ravel()[0]) - sk.log_marginal_likelihood_value_) < 1e-5 * abs(sk.log_marginal_likelihood_value_)
    # predictive (latent) variance k** - k*^T W k*
    Ks = kern.K(Xs, X)
    var = sf2 - np.einsum('ij,jk,ik->i', Ks, post.woodbury_inv, Ks)
    assert np.allclose(var, sd_sk ** 2, atol=2e-6)
    # gradient of the posterior mean against central differences of sklearn's predictor
    G = op.gradients_faithful(Xs, X, ell, sf2, alpha)
    h = 1e-5
    for q in range(d):
        e = np.zeros(d); e[q] = h
        fd = (sk.predict(Xs + e) - sk.predict(Xs - e)) / (2 * h)
        assert np.allclose(G[:, q], fd, atol=5e-6)


def test_eigh_of_gram_equals_svd():
    w = _small()
    _, post, _, _ = _fit_fixed(w)
    alpha = post.woodbury_vector[:, 0]
    G = op.gradients_chunked(w['X'], w['Z'], w['ell'], w['sf2'], alpha)
    C = op.grad_gram_chunked(w['X'], w['Z'], w['ell'], w['sf2'], alpha, chunk=100)
    comps, lam, ratio = op.edr_from_gram(C, 3)
    Vh, S2, r2 = op.svd_faithful(G, 3)
    assert np.allclose(lam, S2, rtol=1e-10)
    assert np.allclose(ratio, r2, rtol=1e-10)
    assert op.principal_angle(comps[:2], Vh[:2]) < 1e-6


def test_vfe_gradients_match_finite_differences():
    rng = np.random.RandomState(5)
    X = rng.standard_normal((60, 3)); y = np.tanh(X[:, 0] - X[:, 1]) + 0.05 * rng.standard_normal(60)
    m = gpy.SparseGPRegression(X, y[:, None], kernel=gpy.RBF(3, ARD=True), num_inducing=8,
                               normalizer=True)
    x0 = m._get_optimizer_array().copy()
    f0, g0 = m._objective_grads(x0)
    idx = rng.choice(x0.size, 10, replace=False)
    for j in idx:
        e = np.zeros_like(x0); e[j] = 1e-6
        fp, _ = m._objective_grads(x0 + e); fm, _ = m._objective_grads(x0 - e)
        fd = (fp - fm) / 2e-6
        assert abs(fd - g0[j]) < 1e-5 * max(1.0, abs(g0[j])), (j, fd, g0[j])


def test_reference_sparse_regression_pin():
    """edrgp/tests/test_edr.py:33-50 with the oracle standing in for GPy."""
    np.random.seed(101)
    N = 50
    noise_var = 0.05
    X = np.linspace(0, 10, 50)[:, None]
    k = gpy.RBF(1)
    y = np.random.multivariate_normal(np.zeros(N), k.K(X) + np.eye(N) * np.sqrt(noise_var)).reshape(-1, 1)
    gp = GaussianProcessRegressor()
    gp.fit(X, y.ravel())
    sgp = SparseGaussianProcessRegressor(num_inducing=12)
    sgp.fit(X, y.ravel())
    assert abs(gp.estimator_.log_likelihood() - sgp.estimator_.log_likelihood()[0][0]) < 0.5


# ---- the reference's EDR tests through the UNMODIFIED reference L3 -------------------------------
def _get_2d_data(ds, mean=None):
    if mean is None:
        mean = [0, 0]
    X = ds.get_gaussian_inputs(eig_values=[1, 0.3], sample_size=500,
                               eig_vectors=np.array([[1, 1], [-1, 1]]), mean=mean)
    y = ds.get_tanh_targets(X, [0.5, 0.5])
    return X, y


@pytest.mark.reference
@pytest.mark.parametrize("mean", [[0, 0], [10, -10]])
def test_reference_mi(reference_edrgp, mean):
    from edrgp.edr import EffectiveDimensionalityReduction
    from edrgp.utils import SVDTransformer
    from edrgp import datasets as ds
    from sklearn.feature_selection import mutual_info_regression
    np.random.seed(1)
    X, y = _get_2d_data(ds, mean)
    edr = EffectiveDimensionalityReduction(SparseGaussianProcessRegressor(num_inducing=30),
                                           SVDTransformer(), n_components=1, normalize=True)
    edr.fit(X, y)
    assert mutual_info_regression(edr.transform(X), y)[0] > 1


@pytest.mark.reference
def test_reference_scaling(reference_edrgp):
    from edrgp.edr import EffectiveDimensionalityReduction
    from edrgp.utils import SVDTransformer
    from edrgp import datasets as ds
    from sklearn.preprocessing import StandardScaler
    np.random.seed(2)
    X, y = _get_2d_data(ds, [10, -10])
    Z0 = StandardScaler().fit_transform(X)[:30].copy()
    est = SparseGaussianProcessRegressor(Z=Z0)
    edr_sc = EffectiveDimensionalityReduction(est, SVDTransformer(), normalize=True)
    edr_sc.fit(X, y, max_iters=50)
    x1 = edr_sc.transform(X - np.mean(X, axis=0))
    edr = EffectiveDimensionalityReduction(est, SVDTransformer(), normalize=False)
    x2 = edr.fit_transform(StandardScaler().fit_transform(X), y, max_iters=50)
    assert np.allclose(x1, x2)


@pytest.mark.reference
@pytest.mark.parametrize("kw", [dict(n_components=2, step=None, normalize=True),
                                dict(n_components=1, step=2, normalize=True),
                                dict(n_components=None, step=0.97, normalize=False)])
def test_reference_loop_matches_unmodified_reference(reference_edrgp, kw):
    """oracle/reference_loop.py (what the GPU box uses as the orchestration oracle) against the
    UNMODIFIED reference classes driven with the same oracle estimator: bit-identical."""
    from edrgp.edr import EffectiveDimensionalityReduction
    from edrgp.utils import SVDTransformer
    from oracle import reference_loop as rl
    rng = np.random.RandomState(0)
    X = rng.standard_normal((300, 6)) * np.linspace(2, .5, 6) + 1.0
    B = np.linalg.qr(rng.standard_normal((6, 2)))[0]
    y = np.tanh(X.dot(B)).sum(1)
    np.random.seed(3)
    ref = EffectiveDimensionalityReduction(SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=15),
                                           SVDTransformer(), **kw)
    ref.fit(X, y, max_iters=5)
    np.random.seed(3)
    out = rl.fit_reference_style(X, y, num_inducing=15, max_iters=5, **kw)
    assert ref.num_iter == out['num_iter']
    assert np.array_equal(ref.components_, out['components_'])
    assert np.array_equal(ref.subspace_variance_ratio_, out['subspace_variance_ratio_'])
    assert np.array_equal(ref.subspace_gradients_, out['subspace_gradients_'])
    assert np.array_equal(ref._first_gradients_, out['_first_gradients_'])
