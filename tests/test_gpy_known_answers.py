"""GPy-PRODUCED known answers: the only numbers in the reference tree that came out of the real GPy.

``/root/reference/examples/BriefIntro.ipynb`` keeps the printed outputs of its author's run (Python 3.6, real
GPy + the reference as it is).  Its EDR section draws its data under ``np.random.seed(3)`` from NumPy's frozen
legacy generator, fits the reference's ``GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True)``
(GPy ``GPRegression`` + L-BFGS-B) under the reference's ``EffectiveDimensionalityReduction(..., SVDTransformer(),
normalize=False)`` and prints

* cell [29]: ``Discrepancy = 0.135``                               (data of cell [21])
* cell [34]: ``Discrepancy = 0.061``                               (targets of cell [32], B_sparse printed by [33])
* cell [35]: ``np.round(edr.components_[:, :2], decimals=3)``      a 10 x 2 table of the fitted EDR directions

These tests replay those cells on top of the UNMODIFIED reference orchestration layer with the oracle standing in
for GPy, and reproduce every printed digit.  What that pins at the GPy boundary: ``RBF`` (ARD) ``K`` /
``gradients_X`` / ``update_gradients_full``, ``Standardize``, the Logexp-transformed L-BFGS-B driver (it lands in
the local optimum GPy found), ``GP.predictive_gradients``.  The variational sparse model is tied to the same
numbers through its Z = X limit at the fitted hyper-parameters -- on the CPU for the oracle's ``vardtc_inference``
and, ``-m gpu``, for the CUDA estimator through the first pass of the reference loop; the optimisation path
(bound + hyper-parameter gradients + L-BFGS-B) through cell [29] with the inducing inputs held at X.  Three decimals is what the notebook
prints; the 1e-8 statements of the other tests are oracle-vs-CUDA, this file is oracle-vs-GPy.

Not reproduced, and not asserted: cell [30] (iterative fit, printed 0.056, replayed 0.049: eight chained L-BFGS-B
runs whose end points depend on the optimiser build -- SciPy's L-BFGS-B was rewritten since; a start moved by 1e-6 or
another L-BFGS memory scatters the replayed value over 0.046 ... 0.054 while cell [29] stays at 0.13498) and cell [37]
(``SparsePCA`` refit: scikit-learn changed the algorithm's normalisation; printed -0.648 / -0.444 / -0.619, replayed
0.653 / 0.432 / 0.622).  ``scipy.sparse.random(..., random_state=11)`` of cell [32] draws differently today, so
B_sparse is taken from cell [33]'s print (8 decimals).
"""
import warnings

import numpy as np
import pytest

from oracle import gpy_restatement as gpy

pytestmark = pytest.mark.reference

# ---- numbers printed by the notebook (examples/BriefIntro.ipynb) ------------------------------------------
CELL_29_DISCREPANCY = 0.135
CELL_34_DISCREPANCY = 0.061
CELL_33_B_SPARSE = np.array([[1., 0.], [0., 0.], [0., -0.60455427], [0., -0.01585381], [0., 0.], [0., 0.],
                             [0., -0.52123653], [0., -0.60214223], [0., 0.], [0., 0.]])
CELL_35_COMPONENTS = np.array([[0.111, -0.002], [-0.99, 0.016], [0.057, 0.027], [0.003, 0.005], [0.045, 0.107],
                               [0.024, -0.242], [0.036, 0.476], [-0.002, 0.838], [0., 0.], [0., -0.]])
PRINT_TOL = 0.5e-3 + 1e-4        # half a unit of the last printed decimal + slack for a value on a rounding edge


def _cells_21_and_32(edrgp):
    """Cells [21] and [32]: the global legacy generator is seeded once, and nothing between the two cells
    draws from it (GPy's ``optimize`` and the SVD are deterministic)."""
    from edrgp.datasets import get_beta_inputs, get_edr_target
    np.random.seed(3)
    X = get_beta_inputs(200, 10)
    B = np.linalg.qr(np.random.normal(size=(10, 2)))[0]
    y = get_edr_target(X.dot(B), sigma=0.1)
    y_sparse = get_edr_target(X.dot(CELL_33_B_SPARSE), sigma=0.1)
    return X, B, y, y_sparse


def _rows_up_to_sign(ours, printed):
    """A component is defined up to its sign (LAPACK build); rows whose printed entries are all 0.000 carry none."""
    sign = np.sign(np.sum(ours * printed, axis=1))
    sign[sign == 0] = 1.
    return ours * sign[:, None]


def _reference_edr(edrgp, estimator, transformer=None):
    from edrgp.utils import SVDTransformer
    return edrgp.EffectiveDimensionalityReduction(estimator, transformer or SVDTransformer(), normalize=False)


@pytest.fixture(scope='module')
def cell_34_fit():
    """Cell [34] replayed once per module: the dense fit on the B_sparse targets."""
    import sys
    from conftest import reference_path
    ref = reference_path()
    if ref not in sys.path:
        sys.path.append(ref)
    import edrgp
    from oracle.estimator import GaussianProcessRegressor
    X, B, y, y_sparse = _cells_21_and_32(edrgp)
    edr = _reference_edr(edrgp, GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        edr.fit(X, y_sparse)
        # the orchestrator keeps the first pass's gradients but not its fitted model (edrgp/base.py:131-136 clones
        # before every fit and ``_last_fit`` refits on the rotated rows): the same deterministic fit once more
        first = GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True).fit(X, y_sparse)
    assert np.array_equal(first.predict_gradient(X), edr._first_gradients_)
    return edrgp, X, y_sparse, edr, first


def test_cell_29_discrepancy(reference_edrgp):
    from edrgp.utils import discrepancy
    from oracle.estimator import GaussianProcessRegressor
    X, B, y, _ = _cells_21_and_32(reference_edrgp)
    edr = _reference_edr(reference_edrgp, GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        edr.fit(X, y)
    d = discrepancy(B, edr.components_.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_29_DISCREPANCY), d


def test_cells_34_and_35_discrepancy_and_components(cell_34_fit):
    edrgp, X, y_sparse, edr, first = cell_34_fit
    from edrgp.utils import discrepancy
    d = discrepancy(CELL_33_B_SPARSE, edr.components_.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_34_DISCREPANCY), d
    ours = _rows_up_to_sign(edr.components_[:, :2], CELL_35_COMPONENTS)
    assert np.max(np.abs(ours - CELL_35_COMPONENTS)) < PRINT_TOL, np.round(ours, 4)


def _fitted_hyperparameters(first):
    m = first.estimator_
    return float(m.kern.variance), np.array(m.kern.lengthscale, dtype=np.float64), float(m.noise_variance)


def test_sparse_oracle_in_its_dense_limit_reproduces_cell_35(cell_34_fit):
    """Z = X at the hyper-parameters GPy's optimum has: the variational posterior IS the exact one, so the
    restated ``VarDTC`` chain + ``predictive_gradients`` must print the same table."""
    edrgp, X, y_sparse, edr, first = cell_34_fit
    from edrgp.utils import SVDTransformer, discrepancy
    sf2, ell, noise = _fitted_hyperparameters(first)
    sparse = gpy.SparseGPRegression(X, y_sparse[:, None], kernel=gpy.RBF(10, sf2, ell, ARD=True), Z=X.copy(),
                                    normalizer=True)
    sparse.noise_variance = noise
    sparse.parameters_changed()
    G = sparse.predictive_gradients(X)[0][:, :, 0]
    G_dense = edr._first_gradients_
    assert np.max(np.abs(G - G_dense)) < 1e-5 * np.max(np.abs(G_dense))
    comps = SVDTransformer().fit(G).components_
    ours = _rows_up_to_sign(comps[:, :2], CELL_35_COMPONENTS)
    assert np.max(np.abs(ours - CELL_35_COMPONENTS)) < PRINT_TOL
    d = discrepancy(CELL_33_B_SPARSE, comps.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_34_DISCREPANCY)


@pytest.mark.gpu
@pytest.mark.parametrize('transformer,route', [('reference_svd', 'fp64'), ('gram_eigh', 'fp64'),
                                               ('gram_eigh', 'int8x6'), ('gram_eigh', 'tf32x3')])
def test_cuda_estimator_in_its_dense_limit_reproduces_cell_35(cell_34_fit, transformer, route):
    """The CUDA estimator (Z = X, the fitted hyper-parameters held fixed) through the first pass of the reference
    loop prints GPy's table: cross-covariance, statistics, Cholesky chain, posterior-mean gradients and
    (``gram_eigh``) the Gram-form eigensolver against numbers the real GPy produced -- in FP64, with the inducing
    statistics on the INT8 tensor cores (``int8x6``) and in the TF32-split mode (``tf32x3``, both tcgen05)."""
    import edrgp_b200 as eb
    from edrgp_b200 import model as emodel
    edrgp, X, y_sparse, edr, first = cell_34_fit
    from edrgp.utils import SVDTransformer, discrepancy
    sf2, ell, noise = _fitted_hyperparameters(first)
    est = eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(10, sf2, ell, ARD=True), Z=X.copy(), normalizer=True,
                                            method='fixed', noise_var=noise,
                                            precision='tf32x3' if route == 'tf32x3' else 'fp64')
    from edrgp_b200 import ops
    ops.set_stats_mode('int8x6' if route == 'int8x6' else 'fp64')
    try:
        # observed on a B200 (gradients against the DENSE model at GPy's optimum, noise on the 1e-8 floor, i.e.
        # beta = 1e8): fp64 3.4e-9, int8x6 2.7e-7, tf32x3 3.2e-5
        _dense_limit_check(est, X, y_sparse, edr, transformer, {'fp64': 1e-7, 'int8x6': 1e-5, 'tf32x3': 1e-3}[route])
    finally:
        ops.set_stats_mode('fp64')


def _dense_limit_check(est, X, y_sparse, edr, transformer, grad_tol):
    import edrgp_b200 as eb
    from edrgp.utils import SVDTransformer, discrepancy
    # the orchestrator's first pass (edrgp/base.py:131-165: fit, predict_gradient on host rows, transformer.fit):
    # with ``n_components=None`` it alone determines ``components_``; the last fit on the rotated rows has no
    # meaning for an estimator whose inducing points are pinned to the unrotated X
    est.fit(X, y_sparse)
    G = est.predict_gradient(X)
    G_dense = edr._first_gradients_
    print('gradients vs GPy-optimum dense model: %.2e' % (np.max(np.abs(G - G_dense)) / np.max(np.abs(G_dense))))
    assert np.max(np.abs(G - G_dense)) < grad_tol * np.max(np.abs(G_dense))
    tr = eb.GramEighTransformer() if transformer == 'gram_eigh' else SVDTransformer()
    comps = np.asarray(tr.fit(G).components_)
    ours = _rows_up_to_sign(comps[:, :2], CELL_35_COMPONENTS)
    assert np.max(np.abs(ours - CELL_35_COMPONENTS)) < PRINT_TOL, np.round(ours, 4)
    d = discrepancy(CELL_33_B_SPARSE, comps.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_34_DISCREPANCY), d


def test_sparse_oracle_optimised_at_Z_equals_X_lands_on_cell_29(reference_edrgp):
    """The OPTIMISATION path of the variational model against a GPy print: with the inducing inputs held at X the
    bound is the exact marginal likelihood (up to GPy's jitter), and L-BFGS-B from GPy's default start lands on the
    optimum the dense model of cell [29] has -- the same 0.135.  (Cell [34]'s data do not offer this: there the
    dense run ends in a poorer local optimum, noise driven to zero, which the variational path does not visit.)"""
    from edrgp.utils import SVDTransformer, discrepancy
    X, B, y, _ = _cells_21_and_32(reference_edrgp)
    sparse = gpy.SparseGPRegression(X, y[:, None], kernel=gpy.RBF(10, ARD=True), Z=X.copy(), normalizer=True)
    sparse.fix_Z = True
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sparse.optimize(max_iters=1000)
    comps = SVDTransformer().fit(sparse.predictive_gradients(X)[0][:, :, 0]).components_
    d = discrepancy(B, comps.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_29_DISCREPANCY), d


@pytest.mark.gpu
def test_cuda_optimiser_at_Z_equals_X_lands_on_cell_29(reference_edrgp):
    """The same through the CUDA estimator: bound and hyper-parameter gradients on the device, L-BFGS-B on the
    host, from GPy's default start (variance 1, lengthscales 1, noise 1), inducing inputs held at X."""
    import edrgp_b200 as eb
    from edrgp.utils import SVDTransformer, discrepancy
    X, B, y, _ = _cells_21_and_32(reference_edrgp)
    est = eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, Z=X.copy(), normalizer=True, method='fixed')
    est.fit(X, y)
    model = est.estimator_
    model.fix_Z = True
    model.optimize(max_iters=1000)
    comps = SVDTransformer().fit(est.predict_gradient(X)).components_
    d = discrepancy(B, comps.T[:, :2])
    assert "{:.3f}".format(d) == "{:.3f}".format(CELL_29_DISCREPANCY), d
    assert model.optimization_runs[-1][2]['warnflag'] == 0
