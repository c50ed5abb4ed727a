"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): kernel entries, gradients and EDR matrices within 1e-8
relative error in FP64.  The checks below use a max-norm relative error and assert 1e-10 or
better, i.e. two orders inside the stated tolerance."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op          # noqa: E402  (the checker, never the product)


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


SHAPES = [
    (500, 10, 20),      # BASELINE config 1
    (1000, 2, 5),       # tiny d, m < one tile
    (777, 7, 33),       # odd d, ragged n and m
    (4096, 32, 256),    # config 2 shape, reduced n
    (3000, 64, 512),    # config 3 shape, reduced n
    (1500, 48, 100),
    (129, 16, 32),
    (1, 4, 3),          # single row
]


@pytest.mark.parametrize("n,d,m", SHAPES)
def test_kuf_matches_oracle(n, d, m):
    from edrgp_b200 import ops
    w = op.make_workload(max(n, m), d, m, seed=n + d)
    X, y = w['X'][:n], w['y'][:n]
    pack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']))
    K, b = ops.kuf(_dev(X), pack, 1.7, y=_dev(y))
    Kref = op.kuf_faithful(X, w['Z'], w['ell'], 1.7)
    assert K.shape == (n, m)
    assert _relerr(K.cpu().numpy(), Kref) < 1e-12
    assert _relerr(b.cpu().numpy(), Kref.T.dot(y)) < 1e-11


@pytest.mark.parametrize("n,d,m", SHAPES)
def test_gradients_and_gram_match_oracle(n, d, m):
    from edrgp_b200 import ops
    w = op.make_workload(max(n, m), d, m, seed=n + d + 1)
    X = w['X'][:n]
    rng = np.random.RandomState(7)
    alpha = rng.standard_normal(m)
    sf2, scale = 1.3, 0.7
    pack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']), _dev(alpha), sf2 * scale)
    G, C = ops.grad_gram(_dev(X), pack)
    Gref = op.gradients_faithful(X, w['Z'], w['ell'], sf2, alpha, scale)
    assert G.shape == (n, d) and C.shape == (d, d)
    assert _relerr(G.cpu().numpy(), Gref) < 1e-11
    Cref = Gref.T.dot(Gref)
    assert _relerr(C.cpu().numpy(), Cref) < 1e-11
    # gradients never leaving the chip give the same Gram matrix
    _, C2 = ops.grad_gram(_dev(X), pack, want_G=False)
    assert torch.equal(C, C2)


def test_coincident_points_are_zeroed_like_gpy():
    """x_i == z_j exactly: GPy's _inv_dist drops the pair; so must the kernel (no NaN, same G)."""
    from edrgp_b200 import ops
    w = op.make_workload(600, 8, 40, seed=11)
    X = w['X'].copy()
    X[:40] = w['Z']                       # first rows coincide with the inducing points
    alpha = np.random.RandomState(1).standard_normal(40)
    pack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']), _dev(alpha), 1.0)
    G, _ = ops.grad_gram(_dev(X), pack, want_C=False)
    Gref = op.gradients_faithful(X, w['Z'], w['ell'], 1.0, alpha)
    assert torch.isfinite(G).all()
    assert _relerr(G.cpu().numpy(), Gref) < 1e-11


def test_far_points_underflow_cleanly():
    from edrgp_b200 import ops
    rng = np.random.RandomState(0)
    X = rng.standard_normal((300, 6)) * 50.0
    Z = rng.standard_normal((16, 6))
    ell = np.full(6, 0.5)
    pack = ops.InducingPack(_dev(Z), _dev(ell))
    K, _ = ops.kuf(_dev(X), pack, 1.0)
    Kref = op.kuf_faithful(X, Z, ell, 1.0)
    assert torch.isfinite(K).all()
    assert np.max(np.abs(K.cpu().numpy() - Kref)) < 1e-300 + 1e-12 * np.max(Kref)


def test_full_size_linearity_property():
    """At a BASELINE-sized row count the oracle is too slow; use linearity in alpha instead:
    G(a1 + a2) == G(a1) + G(a2) and C is the Gram matrix of the G the kernel wrote."""
    from edrgp_b200 import ops
    n, d, m = 200_000, 64, 512
    g = torch.Generator(device='cuda').manual_seed(5)
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    Z = X[torch.randperm(n, device='cuda', generator=g)[:m]].contiguous()
    ell = torch.full((d,), 8.0, dtype=torch.float64, device='cuda') * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
    a1 = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
    a2 = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
    pack = ops.InducingPack(Z, ell, a1, 1.0)
    G1, _ = ops.grad_gram(X, pack, want_C=False)
    G2, _ = ops.grad_gram(X, pack.set_coef(a2, 1.0), want_C=False)
    G12, C12 = ops.grad_gram(X, pack.set_coef(a1 + a2, 1.0))
    err = (G12 - (G1 + G2)).abs().max() / G12.abs().max()
    assert float(err) < 1e-12
    Cref = G12.T @ G12
    assert float((C12 - Cref).abs().max() / Cref.abs().max()) < 1e-12


@pytest.mark.parametrize("n,d,m", SHAPES)
def test_cached_kfu_gradients_match_recompute_and_oracle(n, d, m):
    """The gradient kernel fed from the stored Kfu block (statistics pass) instead of recomputing it."""
    from edrgp_b200 import ops
    w = op.make_workload(max(n, m), d, m, seed=n + d + 2)
    X = w['X'][:n].copy()
    if n > m:
        X[:3] = w['Z'][:3]                  # a few coincident pairs: entries equal to the variance are dropped
    rng = np.random.RandomState(3)
    alpha = rng.standard_normal(m)
    sf2, scale = 1.3, 0.7
    plain = ops.InducingPack(_dev(w['Z']), _dev(w['ell']))
    K, _ = ops.kuf(_dev(X), plain, sf2)
    ldk = m + (m & 1)
    Kbuf = torch.zeros(n, ldk, dtype=torch.float64, device='cuda')
    Kbuf[:, :m] = K
    cpack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']), _dev(alpha), scale)
    G, C = ops.grad_gram_cached(_dev(X), Kbuf, cpack, sf2)
    rpack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']), _dev(alpha), sf2 * scale)
    G2, C2 = ops.grad_gram(_dev(X), rpack)
    Gref = op.gradients_faithful(X, w['Z'], w['ell'], sf2, alpha, scale)
    assert _relerr(G.cpu().numpy(), Gref) < 1e-11
    assert _relerr(C.cpu().numpy(), Gref.T.dot(Gref)) < 1e-11
    assert float((G - G2).abs().max()) <= 1e-13 * float(G2.abs().max())
    _, C3 = ops.grad_gram_cached(_dev(X), Kbuf, cpack, sf2, want_G=False)
    assert torch.equal(C, C3)


@pytest.mark.parametrize("n,d,k", [(1000, 10, 3), (5000, 64, 61), (3001, 33, 20), (2000, 128, 128), (777, 200, 40), (64, 8, 9)])
def test_projection_matches_numpy(n, d, k):
    """EDR.transform: the streaming kernel (few components / wide inputs) and the tensor-pipe path."""
    from edrgp_b200 import ops
    rng = np.random.RandomState(n + k)
    X = rng.standard_normal((n, d)) * 3.0 + 1.0
    V = rng.standard_normal((k, d))
    out = ops.project(_dev(X), _dev(V)).cpu().numpy()
    ref = X.dot(V.T)
    assert out.shape == (n, k)
    assert _relerr(out, ref) < 1e-13


@pytest.mark.parametrize("n,d", [(1, 1), (7, 1), (100000, 1), (500001, 1), (999, 10), (4097, 33), (20000, 64), (3000, 200),
                                 (513, 512), (300, 513), (257, 1300)])
def test_col_moments_match_numpy(n, d):
    """Column sums and sums of squares about a shift, optionally weighted (the scaler / normaliser moments
    and the lengthscale-gradient moment): tree reductions, deterministic."""
    from edrgp_b200 import ops
    rng = np.random.RandomState(n + d)
    X = rng.standard_normal((n, d)) * 3.0 + 1.5
    shift = rng.standard_normal(d)
    w = rng.uniform(0.5, 1.5, size=n)
    Xd = _dev(X if d > 1 else X[:, 0])
    s1, s2 = ops.col_moments(Xd)
    assert np.allclose(s1.cpu().numpy(), X.sum(0), rtol=1e-12, atol=1e-9)
    assert np.allclose(s2.cpu().numpy(), (X ** 2).sum(0), rtol=1e-12)
    s1, s2 = ops.col_moments(Xd, shift=_dev(shift), weight=_dev(w))
    ref1 = (w[:, None] * (X - shift)).sum(0)
    ref2 = (w[:, None] * (X - shift) ** 2).sum(0)
    assert np.allclose(s1.cpu().numpy(), ref1, rtol=1e-12, atol=1e-9)
    assert np.allclose(s2.cpu().numpy(), ref2, rtol=1e-12)
    t1, t2 = ops.col_moments(Xd, shift=_dev(shift), weight=_dev(w))
    assert torch.equal(s1, t1) and torch.equal(s2, t2)


# the plain statistics pass (entries stored, no targets, no row sums): the software-pipelined SIMPLE instantiation of
# kuf_kernel at d in 49..64 / 113..128, the generic one elsewhere -- the path the composite sweep takes
PLAIN_SHAPES = [(3000, 64, 512), (1, 50, 64), (255, 64, 64), (257, 64, 128), (5000, 49, 100), (2049, 60, 448),
                (40000, 64, 512), (300, 64, 96), (513, 32, 128), (1000, 128, 70)]


@pytest.mark.parametrize("n,d,m", PLAIN_SHAPES)
def test_plain_kuf_matches_oracle_with_clip_and_underflow(n, d, m):
    from edrgp_b200 import ops
    w = op.make_workload(max(n, m), d, m, seed=n + d + 5)
    X = w['X'][:n]
    if n > 100:
        X[17] = w['Z'][3]                   # a coincident pair: the clip at r^2 = 0 stores exactly sf2
        X[23] = 1e3 * X[23]                 # a far row: clean underflow to zero
    pack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']))
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    K, _ = ops.kuf(_dev(X), pack, 1.7, flag=flag)
    assert int(flag.item()) == 0
    assert K.shape == (n, m)
    Kref = op.kuf_faithful(X, w['Z'], w['ell'], 1.7)
    assert _relerr(K.cpu().numpy(), Kref) < 1e-12
    if n > 100:
        # exactly sf2 when the rounded r^2 comes out <= 0 (the clip), one or two ulps below when it comes out at +1e-16
        assert abs(float(K[17, 3]) - 1.7) <= 2e-15 and float(K[17, 3]) <= 1.7
        assert float(K[23].max()) == 0.0
    # the instantiation with the row sums (posterior mean) stores bit-identical entries
    alpha = np.random.RandomState(1).standard_normal(m)
    cpack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']), _dev(alpha), 1.0)
    K2, _, mu = ops.kuf(_dev(X), cpack, 1.7, want_mu=True)
    assert torch.equal(K2, K)
    assert _relerr(mu.cpu().numpy(), Kref.dot(alpha)) < 1e-11


def test_plain_kuf_flags_nonfinite_rows_and_leaves_padding_alone():
    from edrgp_b200 import ops
    n, d, m = 1000, 64, 126                 # ldk = 126, four tiles: the last pair of columns is cut by j < m
    w = op.make_workload(n, d, m, seed=3)
    pack = ops.InducingPack(_dev(w['Z']), _dev(w['ell']))
    out = torch.full((n + 5, m), -7.0, dtype=torch.float64, device='cuda')
    K, _ = ops.kuf(_dev(w['X']), pack, 0.9, out=out[:n])
    assert _relerr(K.cpu().numpy(), op.kuf_faithful(w['X'], w['Z'], w['ell'], 0.9)) < 1e-12
    assert bool((out[n:] == -7.0).all())    # rows past n are never written
    X = w['X'].copy()
    X[777, 5] = np.nan
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    ops.kuf(_dev(X), pack, 0.9, flag=flag)
    assert int(flag.item()) == 1
