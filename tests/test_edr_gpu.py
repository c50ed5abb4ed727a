"""End-to-end parity of ``edrgp_b200.EffectiveDimensionalityReduction`` against the oracle chain
(oracle estimator + economy SVD driven through a restatement of the reference iteration logic)
at fixed hyper-parameters, plus the behavioural tests of edrgp/tests/test_edr.py re-run on the
B200 classes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op                  # noqa: E402
from oracle import reference_loop as oloop          # noqa: E402


def _make(n, d, k_true, seed):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n, d)) * np.linspace(2.0, 0.5, d) + rng.standard_normal(d)
    B = np.linalg.qr(rng.standard_normal((d, k_true)))[0]
    y = np.tanh(X.dot(B)).sum(1) + 0.05 * rng.standard_normal(n)
    return X, y, B


CASES = [
    dict(n=500, d=10, m=20, k=2, step=None, normalize=True),       # BASELINE config 1 shape
    dict(n=800, d=8, m=30, k=3, step=None, normalize=False),
    dict(n=600, d=6, m=25, k=1, step=2, normalize=True),            # iterative shrinking
    dict(n=600, d=6, m=25, k=None, step=0.97, normalize=True),      # adaptive step
    dict(n=700, d=9, m=20, k=None, step=None, normalize=True),      # k = d: single pass
]


@pytest.mark.parametrize("case", CASES)
def test_edr_fit_matches_oracle_chain(case):
    import edrgp_b200 as eb
    X, y, _ = _make(case['n'], case['d'], 2, seed=case['n'] + case['d'])
    kw = dict(n_components=case['k'], step=case['step'], normalize=case['normalize'])
    np.random.seed(7)
    ref = oloop.fit_reference_style(X, y, num_inducing=case['m'], max_iters=0, **kw)
    np.random.seed(7)
    edr = eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=case['m']),
        eb.GramEighTransformer(), **kw)
    edr.fit(X, y, max_iters=0)
    assert edr.num_iter == ref['num_iter']
    assert edr.components_.shape == ref['components_'].shape
    assert op.principal_angle(edr.components_, ref['components_']) < 1e-6
    for a, b in zip(edr.components_, ref['components_']):
        assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-7 * np.abs(b).max()
    assert np.allclose(edr.subspace_variance_ratio_, ref['subspace_variance_ratio_'], rtol=1e-7, atol=1e-12)
    assert np.allclose(edr.subspace_variance_, ref['subspace_variance_'], rtol=1e-7)
    assert np.allclose(edr._first_gradients_, ref['_first_gradients_'], rtol=0, atol=1e-8 * np.abs(ref['_first_gradients_']).max())
    Xt = edr.transform(X[:20])
    assert np.allclose(np.abs(Xt), np.abs(X[:20].dot(ref['components_'].T)), rtol=1e-6, atol=1e-9)
    assert np.allclose(np.abs(edr.subspace_gradients_), np.abs(ref['subspace_gradients_']), rtol=1e-5, atol=1e-8)


def _two_d(n=500, seed=0):
    rng = np.random.RandomState(seed)
    mean = np.array([1.0, -2.0])
    cov = np.array([[1.0, 0.6], [0.6, 2.0]])
    X = rng.multivariate_normal(mean, cov, n)
    y = np.tanh((X - mean).dot([0.5, -0.5]))
    return X, y


def _edr(k=1, m=40, **kw):
    import edrgp_b200 as eb
    return eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=m),
        eb.GramEighTransformer(), n_components=k, **kw)


def test_mi_like_reference():
    """edrgp/tests/test_edr.py:53-61: the EDR component carries the target (MI > 1)."""
    from sklearn.feature_selection import mutual_info_regression
    X, y = _two_d()
    np.random.seed(0)
    edr = _edr(normalize=True).fit(X, y, max_iters=200)
    mi = mutual_info_regression(edr.transform(X), y, random_state=0)[0]
    assert mi > 1


def test_translation_invariance_like_reference():
    """edrgp/tests/test_edr.py:64-77: components_ unchanged under a shift of X (rtol 1e-3)."""
    X, y = _two_d()
    np.random.seed(0)
    a = _edr(normalize=True).fit(X, y, max_iters=200)
    np.random.seed(0)
    b = _edr(normalize=True).fit(X + np.array([10.0, -3.0]), y, max_iters=200)
    assert np.allclose(a.components_, b.components_, rtol=1e-3)


def test_scaling_equivalence_like_reference():
    """edrgp/tests/test_edr.py:103-117: normalize=True on raw X == normalize=False on standardised X."""
    from sklearn.preprocessing import StandardScaler
    X, y = _two_d()
    np.random.seed(0)
    a = _edr(normalize=True).fit(X, y, max_iters=100)
    Xs = StandardScaler().fit_transform(X)
    np.random.seed(0)
    b = _edr(normalize=False).fit(Xs, y, max_iters=100)
    assert np.allclose(a.transform(X) - a.transform(X).mean(0), b.transform(Xs) - b.transform(Xs).mean(0), atol=1e-6)


def test_preprocessor_chain_like_reference():
    """edrgp/tests/test_edr.py:80-100: PCA preprocessor on data with two near-null directions;
    components live in raw feature space and are shift invariant."""
    from sklearn.decomposition import PCA
    rng = np.random.RandomState(1)
    X2, y = _two_d(seed=1)
    X = np.hstack([X2, 1e-3 * rng.standard_normal((X2.shape[0], 2))])
    import edrgp_b200 as eb
    def make():
        return eb.EffectiveDimensionalityReduction(
            eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=40),
            eb.GramEighTransformer(), n_components=1, normalize=True, preprocessor=PCA(n_components=2))
    np.random.seed(0)
    a = make().fit(X, y, max_iters=100)
    np.random.seed(0)
    b = make().fit(X + 5.0, y, max_iters=100)
    assert a.components_.shape == (1, 4)
    assert np.allclose(a.components_, b.components_, rtol=1e-3, atol=1e-6)


def test_refit_with_host_transformer():
    """examples/sPCAvsPCA.ipynb cell 12: refit(SparsePCA) on the gradients kept at fit time."""
    from sklearn.decomposition import SparsePCA
    X, y, _ = _make(400, 6, 2, seed=3)
    np.random.seed(0)
    edr = _edr(k=2, m=25).fit(X, y, max_iters=0)
    edr.refit(SparsePCA(n_components=2, random_state=0))
    assert edr.refit_components_.shape[1] == 6
    assert edr.transform(X, refitted=True).shape == (400, edr.refit_components_.shape[0])
    assert edr.refit_subspace_variance_ratio_.shape[0] >= 1


def test_get_estimator_gradients_and_importances():
    X, y, _ = _make(400, 6, 2, seed=5)
    np.random.seed(0)
    edr = _edr(k=2, m=25).fit(X, y, max_iters=0)
    g = edr.get_estimator_gradients(X[:30])
    assert g.shape == (30, 6) and np.all(np.isfinite(g))
    assert edr.feature_importances_.shape == (2, 6)
    assert edr.inverse_transform(edr.transform(X[:5])).shape == (5, 6)


def test_device_pca_matches_sklearn_and_chains():
    """the PCA preprocessor on the device (examples/chain_PCA-EDRGP.ipynb): same
    components / variances as sklearn's PCA, and the chained EDR fit equals the host-PCA chain."""
    from sklearn.decomposition import PCA
    import edrgp_b200 as eb
    rng = np.random.RandomState(4)
    X = rng.standard_normal((3000, 9)).dot(rng.standard_normal((9, 9))) + rng.standard_normal(9) * 3
    ref = PCA(n_components=4, svd_solver='full').fit(X)
    dev = eb.DevicePCA(n_components=4)
    Xt = dev.fit_transform(X)
    assert np.allclose(dev.explained_variance_, ref.explained_variance_, rtol=1e-10)
    assert np.allclose(dev.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-10)
    assert np.allclose(dev.mean_, ref.mean_, rtol=1e-12, atol=1e-12)
    for a, b in zip(dev.components_, ref.components_):
        assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-9
    assert np.allclose(np.abs(Xt), np.abs(ref.transform(X)), rtol=1e-8, atol=1e-9)
    assert np.allclose(dev.transform(X[:7]), Xt[:7], rtol=1e-10, atol=1e-10)

    X2, y = _two_d(seed=1)
    Xw = np.hstack([X2, 1e-3 * rng.standard_normal((X2.shape[0], 2))])

    def make(pre):
        return eb.EffectiveDimensionalityReduction(
            eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=40),
            eb.GramEighTransformer(), n_components=1, normalize=True, preprocessor=pre)
    np.random.seed(0)
    a = make(PCA(n_components=2, svd_solver='full')).fit(Xw, y, max_iters=0)
    np.random.seed(0)
    b = make(eb.DevicePCA(n_components=2)).fit(Xw, y, max_iters=0)
    assert a.components_.shape == b.components_.shape == (1, 4)
    assert op.principal_angle(a.components_, b.components_) < 1e-6


def test_edr_rejects_bad_input():
    X, y, _ = _make(200, 5, 2, seed=9)
    Xb = X.copy(); Xb[3, 2] = np.nan
    with pytest.raises(ValueError):
        _edr(k=2, m=10).fit(Xb, y, max_iters=0)
    yb = y.copy(); yb[5] = np.inf
    with pytest.raises(ValueError):
        _edr(k=2, m=10).fit(X, yb, max_iters=0)
    with pytest.raises(ValueError):
        _edr(k=2, m=10).fit(X, y[:-1], max_iters=0)
    with pytest.raises(ValueError):
        _edr(k=2, m=10).fit(X[:, 0], y, max_iters=0)
