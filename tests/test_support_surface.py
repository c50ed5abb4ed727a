"""The support surface a user of the reference imports next to the estimator (SURVEY section 2 "support" row:
``edrgp/utils.py``, ``edrgp/datasets.py``): generators bit-identical to the reference's under the same seed,
``ort_space`` / ``discrepancy`` against the reference's, ``subspace_variance_ratio`` (row form, Gram matrix on the
device) against the reference's on the GPU."""
import numpy as np
import pytest

import edrgp_b200 as eb
from edrgp_b200 import datasets as ds


@pytest.mark.reference
def test_generators_draw_what_the_reference_draws(reference_edrgp):
    from edrgp import datasets as ref
    calls = [
        lambda m: m.get_beta_inputs(50, 7),
        lambda m: m.get_beta_inputs(20, 3, tau=2.5),
        lambda m: m.get_gaussian_inputs(40, [1, 0.3], eig_vectors=np.array([[1, 1], [-1, 1]])),
        lambda m: m.get_gaussian_inputs(30, [2., 1., .5], mean=np.array([1., -1., 0.])),
        lambda m: m.get_gaussian_inputs(10, [2., 1., .5, .1]),
        lambda m: m.get_tanh_targets(m.get_beta_inputs(25, 4), [0.5, 0.5, -1., 2.], bias=0.1, noise_std=0.2),
        lambda m: m.get_edr_target(m.get_beta_inputs(25, 1), sigma=0.1),
        lambda m: m.get_edr_target(m.get_beta_inputs(25, 2), sigma=0.1),
        lambda m: m.get_edr_target(m.get_beta_inputs(25, 3)),
        lambda m: m.get_branin_targets((m.get_beta_inputs(25, 2) + 1) / 2, noise_std=0.3),
        lambda m: m.get_branin_targets((m.get_beta_inputs(25, 2) + 1) / 2),
    ]
    for i, call in enumerate(calls):
        np.random.seed(100 + i)
        want = call(ref)
        np.random.seed(100 + i)
        got = call(ds)
        assert got.shape == want.shape and got.dtype == want.dtype, i
        assert np.array_equal(got, want), i
    with pytest.raises(ValueError):
        ds.get_gaussian_inputs(5, [1, 2], eig_vectors=np.eye(3))
    with pytest.raises(ValueError):
        ds.get_tanh_targets(np.zeros((4, 3)), [1., 2.])


@pytest.mark.reference
def test_ort_space_and_discrepancy_like_the_reference(reference_edrgp):
    from edrgp import utils as ref
    rng = np.random.RandomState(0)
    for shape in ((6, 2), (5, 5), (8, 1)):
        A = rng.normal(size=shape)
        U, Uref = eb.ort_space(A), ref.ort_space(A)
        assert U.shape == Uref.shape == (shape[0], shape[0] - shape[1])
        if U.shape[1]:
            assert np.allclose(U.T.dot(A), 0, atol=1e-12) and np.allclose(U.T.dot(U), np.eye(U.shape[1]), atol=1e-12)
            assert np.allclose(U.dot(U.T), Uref.dot(Uref.T), atol=1e-12)
    A = rng.normal(size=(6, 2))
    A2 = np.c_[A, A[:, :1] * 2.]                     # rank 2 in three columns
    assert eb.ort_space(A2).shape == ref.ort_space(A2).shape == (6, 4)
    B, V = np.linalg.qr(rng.normal(size=(7, 2)))[0], np.linalg.qr(rng.normal(size=(7, 3)))[0]
    assert eb.discrepancy(B, V) == pytest.approx(ref.discrepancy(B, V), rel=1e-14)
    assert eb.SVDTransformer is eb.GramEighTransformer


@pytest.mark.gpu
@pytest.mark.reference
def test_row_form_variance_ratio_matches_the_reference(reference_edrgp):
    import torch
    from edrgp import utils as ref
    rng = np.random.RandomState(1)
    G = rng.normal(size=(3001, 9)) * np.linspace(3., .1, 9)
    V = np.linalg.qr(rng.normal(size=(9, 4)))[0]
    for W in (V, V.dot(np.array([[1., .3, 0, 0], [0, 1., 0, 0], [0, 0, 2., 0], [0, 0, .5, 1.]]))):   # orthonormal / not
        var, ratio = eb.subspace_variance_ratio(G, W)
        var_ref, ratio_ref = ref.subspace_variance_ratio(G, W)
        assert np.allclose(var, var_ref, rtol=1e-12) and np.allclose(ratio, ratio_ref, rtol=1e-12)
    var_dev, ratio_dev = eb.subspace_variance_ratio(torch.as_tensor(G, device='cuda'), V)
    assert np.allclose(ratio_dev, ref.subspace_variance_ratio(G, V)[1], rtol=1e-12)
