"""The plug-in swap of INTEGRATION.md section 1, executed: the UNMODIFIED reference orchestrator
(``edrgp.EffectiveDimensionalityReduction``, edrgp/edr.py:11 over edrgp/base.py:435-466 -- from
/root/reference, or on the GPU box from the byte-identical copy ``oracle/build_ref.py`` stages under
``oracle/_ref/``) drives the CUDA estimator and transformer with HOST arrays, exactly as a user of the
reference would after changing two imports.  Everything is compared with the golden fixtures, which
were produced by the same unmodified orchestrator on top of the oracle estimator and the reference's
own ``SVDTransformer`` (tests/golden/make_golden.py).

Also here: value tests of the repo's own ``EffectiveDimensionalityReduction.refit`` /
``get_estimator_gradients`` / ``transform`` / ``feature_importances_`` against the reference's
(edrgp/base.py:202-239, edrgp/edr.py:115-140,199-289) through the same fixtures.
"""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import pipeline as op

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.npz')))
FIXED = [p for p in GOLDEN if 'optimised' not in p]
IDS = [os.path.basename(p)[:-4] for p in FIXED]


def _cfg(g):
    c = {k[4:]: g[k].item() for k in g.files if k.startswith('cfg_')}
    for k in ('k', 'step'):
        if c[k] == -1:
            c[k] = None
    if c['step'] is not None and float(c['step']) >= 1:
        c['step'] = int(c['step'])
    if c['k'] is not None:
        c['k'] = int(c['k'])
    return c


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def _same_up_to_row_signs(a, b, tol):
    """Components are defined up to the sign of each row (SVD vs eigh conventions)."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    assert a.shape == b.shape
    sign = np.sign(np.sum(a * b, axis=1))
    return _rel(a * sign[:, None], b) < tol


def _check_fit(edr, g, c):
    assert edr.num_iter == int(g['num_iter'])
    assert op.principal_angle(edr.components_, g['components_']) < 1e-6
    assert np.allclose(edr.subspace_variance_ratio_, g['subspace_variance_ratio_'], rtol=1e-7, atol=1e-12)
    assert np.allclose(edr.subspace_variance_, g['subspace_variance_'], rtol=1e-7, atol=1e-12)
    assert _rel(edr._first_gradients_, g['first_gradients']) < 1e-8
    ll = float(edr.estimator_.estimator_.log_likelihood()[0, 0])
    assert abs(ll - float(g['final_loglik'])) < 1e-8 * abs(float(g['final_loglik']))
    # the projected basis is fixed only up to the sign of each row (SVD vs eigh conventions): compare what does
    # not depend on it -- gradients mapped back to the raw features, Gram forms of projection and importances
    Xq = g['X'][:40] + 0.01
    assert _rel(edr.get_estimator_gradients(Xq), g['get_estimator_gradients']) < 1e-6
    T, F = edr.transform(Xq), edr.feature_importances_
    assert T.shape == g['transform'].shape and F.shape == g['feature_importances_'].shape
    assert _rel(T.dot(T.T), g['transform'].dot(g['transform'].T)) < 1e-6
    assert _rel(F.T.dot(F), g['feature_importances_'].T.dot(g['feature_importances_'])) < 1e-6
    if c['k'] is not None:                  # well separated leading directions: row by row, up to sign
        assert _same_up_to_row_signs(edr.components_, g['components_'], 1e-6)


def _check_refit(edr, g, c, make_svd):
    from sklearn.decomposition import SparsePCA
    kk = 2 if c['k'] is None else c['k']
    Xq = g['X'][:40] + 0.01
    edr.refit(make_svd(kk))
    assert _same_up_to_row_signs(edr.refit_components_, g['refit_components_'], 1e-6)
    assert np.allclose(edr.refit_subspace_variance_, g['refit_subspace_variance_'], rtol=1e-7)
    assert np.allclose(edr.refit_subspace_variance_ratio_, g['refit_subspace_variance_ratio_'], rtol=1e-7)
    Tr = edr.transform(Xq, refitted=True)
    assert _rel(Tr.dot(Tr.T), g['refit_transform'].dot(g['refit_transform'].T)) < 1e-6
    edr.refit(make_svd(kk), g['refit_rows'])
    assert _same_up_to_row_signs(edr.refit_components_, g['refit_rows_components_'], 1e-6)
    assert np.allclose(edr.refit_subspace_variance_ratio_, g['refit_rows_subspace_variance_ratio_'], rtol=1e-7)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)          # all-zero sparse components are dropped
        edr.refit(SparsePCA(n_components=kk, alpha=0.5, random_state=0))
    # a host transformer (coordinate descent) fed with gradients that agree to 1e-9
    assert edr.refit_components_.shape == g['refit_sparse_components_'].shape
    assert _same_up_to_row_signs(edr.refit_components_, g['refit_sparse_components_'], 1e-4)
    assert np.allclose(edr.refit_subspace_variance_ratio_, g['refit_sparse_subspace_variance_ratio_'], rtol=1e-4)


@pytest.mark.reference
@pytest.mark.parametrize("path", FIXED, ids=IDS)
def test_unmodified_reference_orchestrator_with_cuda_plugins(reference_edrgp, path):
    """INTEGRATION.md section 1 as written: reference L3 + CUDA estimator + CUDA transformer."""
    import edrgp_b200 as eb
    from edrgp.edr import EffectiveDimensionalityReduction as ReferenceEDR
    from edrgp.utils import SVDTransformer as ReferenceSVD
    assert ReferenceEDR.__module__ == 'edrgp.edr' and 'edrgp_b200' not in ReferenceEDR.__module__
    g = np.load(path)
    c = _cfg(g)
    np.random.seed(100 + int(c['seed']))
    edr = ReferenceEDR(eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=int(c['m'])),
                       eb.GramEighTransformer(), n_components=c['k'], step=c['step'], normalize=bool(c['normalize']))
    edr.fit(g['X'], g['y'], max_iters=int(c['max_iters']))
    assert isinstance(edr._first_gradients_, np.ndarray)           # host arrays crossed the plug-in boundary
    assert type(edr.transformer_).__name__ == 'GramEighTransformer'
    _check_fit(edr, g, c)
    if c['k'] is not None:
        assert _rel(np.abs(edr.subspace_gradients_), np.abs(g['subspace_gradients_'])) < 1e-6
    _check_refit(edr, g, c, lambda k: eb.GramEighTransformer(n_components=k))
    _check_refit(edr, g, c, lambda k: ReferenceSVD(n_components=k))


@pytest.mark.reference
def test_unmodified_reference_orchestrator_optimises_with_cuda_estimator(reference_edrgp):
    """The default ``method='optimize'`` path under the reference L3 (opt_kws flow through unchanged,
    edrgp/gp_model/base.py:67-69): L-BFGS trajectories are not bit-comparable, the fitted subspace is."""
    import edrgp_b200 as eb
    from edrgp.edr import EffectiveDimensionalityReduction as ReferenceEDR
    path = [p for p in GOLDEN if 'optimised' in p][0]
    g = np.load(path)
    c = _cfg(g)
    np.random.seed(100 + int(c['seed']))
    edr = ReferenceEDR(eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=int(c['m'])),
                       eb.GramEighTransformer(), n_components=c['k'], step=c['step'], normalize=bool(c['normalize']))
    edr.fit(g['X'], g['y'], max_iters=int(c['max_iters']))
    assert edr.num_iter == int(g['num_iter'])
    assert edr.components_.shape == g['components_'].shape
    # 25 L-BFGS iterations from the same start on the same objective: same basin, close subspace
    assert op.principal_angle(edr.components_, g['components_']) < 5e-2
    ll = float(edr.estimator_.estimator_.log_likelihood()[0, 0])
    assert abs(ll - float(g['final_loglik'])) < 2e-2 * abs(float(g['final_loglik']))


@pytest.mark.parametrize("path", FIXED, ids=IDS)
def test_cuda_edr_surface_matches_reference_values(path):
    """The repo's own orchestrator: fit, refit (Gram form, row subset, host transformer),
    get_estimator_gradients, transform, feature_importances_ -- values, not shapes."""
    import edrgp_b200 as eb
    from oracle.reference_loop import EconomySVDTransformer
    g = np.load(path)
    c = _cfg(g)
    np.random.seed(100 + int(c['seed']))
    edr = eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=int(c['m'])), eb.GramEighTransformer(),
        n_components=c['k'], step=c['step'], normalize=bool(c['normalize']))
    edr.fit(g['X'], g['y'], max_iters=int(c['max_iters']))
    _check_fit(edr, g, c)
    _check_refit(edr, g, c, lambda k: eb.GramEighTransformer(n_components=k))
    _check_refit(edr, g, c, lambda k: EconomySVDTransformer(n_components=k))


@pytest.mark.reference
@pytest.mark.parametrize("transformer", ["gram", "host"])
def test_block_edr_matches_reference_block_edr(reference_edrgp, transformer):
    """``BlockEDR`` (eigh on diagonal blocks of the reduced Gram matrix) against the UNMODIFIED reference class
    (edrgp/base.py:520-766) driven with the oracle estimator and the reference's SVDTransformer."""
    import numpy.matlib  # noqa: F401  (the reference calls np.matlib.repmat without importing it)
    import edrgp_b200 as eb
    from edrgp.base import BlockEDR as ReferenceBlockEDR
    from edrgp.utils import SVDTransformer as ReferenceSVD
    from oracle.estimator import SparseGaussianProcessRegressor as OracleSGPR
    from oracle.reference_loop import EconomySVDTransformer
    g = np.load([p for p in GOLDEN if 'small_adaptive' in p][0])
    X, y = g['X'], g['y']                                    # (300, 6)
    blocks, ncomp = [[0, 1, 2], [3, 4, 5]], [1, 2]
    np.random.seed(21)
    ref = ReferenceBlockEDR(OracleSGPR('RBF', {'ARD': True}, num_inducing=15), ReferenceSVD(), n_components=ncomp,
                            blocks=[list(b) for b in blocks])
    ref.fit(X, y, max_iters=0)
    np.random.seed(21)
    tr = eb.GramEighTransformer() if transformer == "gram" else EconomySVDTransformer()
    edr = eb.BlockEDR(eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=15), tr, n_components=ncomp,
                      blocks=[list(b) for b in blocks])
    edr.fit(X, y, max_iters=0)
    assert edr.components_.shape == ref.components_.shape == (3, 6)
    assert np.all(edr.components_[0, 3:] == 0) and np.all(edr.components_[1:, :3] == 0)      # block diagonal
    assert _same_up_to_row_signs(edr.components_, ref.components_, 1e-6)
    assert np.allclose(edr.subspace_variance_ratio_, ref.subspace_variance_ratio_, rtol=1e-7)
    assert np.allclose(edr.subspace_variance_, ref.subspace_variance_, rtol=1e-7)
    assert _rel(edr._first_gradients_, ref._first_gradients_) < 1e-8
    ll, ll_ref = (float(e.estimator_.estimator_.log_likelihood()[0, 0]) for e in (edr, ref))
    assert abs(ll - ll_ref) < 1e-8 * abs(ll_ref)
    # per-block refit
    ref.refit(ReferenceSVD())
    edr.refit(eb.GramEighTransformer() if transformer == "gram" else EconomySVDTransformer())
    assert _same_up_to_row_signs(edr.refit_components_, ref.refit_components_, 1e-6)
    assert np.allclose(edr.refit_subspace_variance_ratio_, ref.refit_subspace_variance_ratio_, rtol=1e-7)
