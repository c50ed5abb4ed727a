"""The exact-product INT8 route of the inducing statistics (edrgp_inducing_stats_i8, `stats='int8x6'`):
P = Kfu^T Kfu from six signed radix-256 digits of K / (4 sf2) on the tcgen05 INT8 tensor cores.

What is checked, through the C ABI:
  * known answers: entries that are exact in 48-bit fixed point give the EXACT integer result (bit for bit);
  * against the FP64 DMMA reduction and a NumPy product on seeded kernels: 3e-14 relative (max norm) for a few
    hundred rows, 5e-15 from 20 000 rows on (the rounding at 2^-49 and the dropped digit pairs are zero-mean, so the
    error falls like 1 / sqrt(n) to FP64 rounding level), exact symmetry, b and y^T y to FP64 rounding, accumulate mode;
  * the composite sweep with the route switched on against the FP64 route and against the CPU oracle, at the
    tolerances of the FP64 path (gradients / EDR matrix 1e-8, BASELINE.json north_star)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op                  # noqa: E402  (the checker, never the product)


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


@pytest.fixture
def int8_route():
    from edrgp_b200 import ops
    ops.set_stats_mode('int8x6')
    yield ops
    ops.set_stats_mode('fp64')


@pytest.mark.parametrize("n,m", [(1, 2), (127, 5), (128, 128), (4097, 130), (100000, 256), (9000, 1024)])
def test_exact_on_fixed_point_entries(n, m):
    """K = j / 2^16 with integer j in [0, 2^16]: the digits hold j exactly, every product and every sum fits 53
    bits, so P must equal the integer result bit for bit (and so must b and y^T y for integer targets)."""
    from edrgp_b200 import ops
    rs = np.random.RandomState(n + m)
    J = rs.randint(0, 2 ** 16 + 1, size=(n, m)).astype(np.float64)
    J[0, 0] = 2.0 ** 16                                   # an entry equal to sf2 (x == z)
    J[-1, -1] = 0.0
    yv = rs.randint(-8, 9, size=n).astype(np.float64)
    ldk = m + (m & 1)
    K = torch.zeros(n, ldk, dtype=torch.float64, device='cuda')
    K[:, :m] = torch.as_tensor(J / 2.0 ** 16, device='cuda')
    P, byy = ops.inducing_stats_i8(K, torch.as_tensor(yv, device='cuda'), 1.0, m)
    P_int = J.T.dot(J)                                    # exact: every partial sum < 2^53
    assert np.array_equal(P.cpu().numpy() * 2.0 ** 32, P_int)
    assert np.array_equal(byy.cpu().numpy()[:m] * 2.0 ** 16, J.T.dot(yv))
    assert float(byy[m]) == float(yv.dot(yv))


@pytest.mark.parametrize("n,d,m,sf2", [(129, 6, 7, 0.3), (1000, 8, 64, 1.0), (20001, 16, 300, 1.7),
                                       (40000, 64, 512, 2.5), (3000, 32, 1024, 1e-3), (2500, 16, 2048, 40.0)])
def test_matches_fp64_reduction(n, d, m, sf2):
    from edrgp_b200 import ops
    rs = np.random.RandomState(m)
    X = rs.standard_normal((n, d))
    Z = X[rs.choice(n, m, replace=False)] if n >= m else rs.standard_normal((m, d))
    ell = np.sqrt(d) * (0.6 + rs.rand(d))
    y = torch.as_tensor(rs.standard_normal(n), device='cuda')
    Xd = torch.as_tensor(X, device='cuda')
    K = torch.empty(n, m + (m & 1), dtype=torch.float64, device='cuda')
    ops.kuf(Xd, ops.InducingPack(torch.as_tensor(Z, device='cuda'), torch.as_tensor(ell, device='cuda')), sf2, out=K)
    P0, b0 = ops.inducing_stats(K, y, m)
    P1, b1 = ops.inducing_stats_i8(K, y, sf2, m)
    Kh = K[:, :m].cpu().numpy()
    tol = 5e-15 if n >= 20000 else 3e-14
    assert _rel(P1.cpu().numpy(), Kh.T.dot(Kh)) < tol
    assert _rel(P1.cpu().numpy(), P0.cpu().numpy()) < tol
    assert float((P1 - P1.T).abs().max()) == 0.0
    assert _rel(b1.cpu().numpy()[:m], b0.cpu().numpy()[:m]) < 1e-13
    assert abs(float(b1[m]) - float(b0[m])) < 1e-14 * float(b0[m])
    # accumulate mode and the call without targets
    P2, b2 = ops.inducing_stats_i8(K, y, sf2, m, P=P1.clone(), b_yy=b1.clone(), accumulate=True)
    assert _rel(P2.cpu().numpy(), 2 * P1.cpu().numpy()) < 1e-15
    assert _rel(b2.cpu().numpy(), 2 * b1.cpu().numpy()) < 1e-15
    P3, none = ops.inducing_stats_i8(K, None, sf2, m)
    assert none is None and torch.equal(P3, P1)


def test_rejects_what_it_does_not_cover():
    from edrgp_b200 import ops, _lib
    K = torch.zeros(64, 2050, dtype=torch.float64, device='cuda')
    with pytest.raises(ValueError):
        ops.inducing_stats_i8(K, None, 1.0, 2050)
    K = torch.zeros(64, 8, dtype=torch.float64, device='cuda')
    with pytest.raises(_lib.EdrgpError):
        ops.inducing_stats_i8(K, None, 0.0, 8)
    with pytest.raises(ValueError):
        ops.set_stats_mode('int4')


@pytest.mark.parametrize("n,d,m,chunk", [(6000, 16, 96, 2048), (10000, 32, 256, 4096), (5000, 64, 512, 1024)])
def test_sweep_with_int8_statistics(n, d, m, chunk, int8_route):
    """The composite sweep (several row blocks, the last one ragged) with the INT8 route against the CPU oracle and
    against the same sweep on the FP64 route."""
    from edrgp_b200 import model
    w = op.make_workload(n, d, m, seed=d + m, k_true=2)
    w['y'] = 1.5 * w['y'] + 0.3
    mean, std = w['y'].mean(), w['y'].std()
    yn = (w['y'] - mean) / std
    P, b, yy = op.inducing_stats_chunked(w['X'], yn, w['Z'], w['ell'], w['sf2'])
    sol = op.solve_from_stats(op.kuu(w['Z'], w['ell'], w['sf2']), P, b, yy, n, w['sf2'], w['noise'])
    G_ref = op.gradients_chunked(w['X'], w['Z'], w['ell'], w['sf2'], sol['alpha'], scale=std)

    def run():
        mod = model.SparseGPRegression(w['X'], w['y'][:, None], kernel=model.RBF(d, w['sf2'], w['ell'], ARD=True), Z=w['Z'],
                                       normalizer=True, noise_var=w['noise'], chunk_rows=chunk)
        Pd, byy = mod._stats
        Pd, byy = Pd.clone(), byy.clone()
        G, C = mod.gradient_gram(want_G=True, want_C=True)
        assert mod._fixed is not None, "the composite sweep did not run"
        return Pd.cpu().numpy(), byy.cpu().numpy(), G.cpu().numpy(), C.cpu().numpy(), float(mod.log_likelihood()[0, 0])

    assert int8_route.get_stats_mode() == 'int8x6'
    P8, b8, G8, C8, ll8 = run()
    assert _rel(P8, P) < 3e-14
    assert _rel(b8[:m], b) < 1e-11
    assert abs(ll8 - sol['bound']) < 1e-9 * abs(sol['bound'])
    assert _rel(G8, G_ref) < 1e-8
    assert _rel(C8, G_ref.T.dot(G_ref)) < 1e-8
    int8_route.set_stats_mode('fp64')
    P64, b64, G64, C64, ll64 = run()
    assert _rel(P8, P64) < 3e-14
    assert _rel(G8, G64) < 1e-8
    assert _rel(C8, C64) < 1e-8


def test_generic_pass_with_int8_statistics(int8_route):
    """d > 64 runs the row passes outside the composite sweep (feature-blocked kernels): the INT8 route there, against
    the FP64 route and the oracle; also the bound and the hyper-parameter gradients (the optimiser's evaluation)."""
    from edrgp_b200 import model
    n, d, m = 3000, 96, 160
    w = op.make_workload(n, d, m, seed=5, k_true=2)

    def run():
        mod = model.SparseGPRegression(w['X'], w['y'][:, None], kernel=model.RBF(d, w['sf2'], w['ell'], ARD=True), Z=w['Z'],
                                       normalizer=True, noise_var=w['noise'], chunk_rows=1024)
        P = mod._stats[0].cpu().numpy().copy()
        G, C = mod.gradient_gram(want_G=True, want_C=True)
        assert mod._fixed is None
        ll = float(mod.log_likelihood()[0, 0])
        mod._need_grad = True
        mod.parameters_changed()
        mod._need_grad = False
        return P, G.cpu().numpy(), ll, np.array(mod.grad_lengthscale), mod.grad_variance, mod.grad_noise

    P8, G8, ll8, gl8, gv8, gn8 = run()
    int8_route.set_stats_mode('fp64')
    P64, G64, ll64, gl64, gv64, gn64 = run()
    assert 0 < _rel(P8, P64) < 3e-14            # the route ran (not bit-identical) and agrees
    assert _rel(G8, G64) < 1e-8
    assert abs(ll8 - ll64) < 1e-9 * abs(ll64)
    assert _rel(gl8, gl64) < 1e-7 and abs(gv8 - gv64) < 1e-7 * abs(gv64) and abs(gn8 - gn64) < 1e-7 * abs(gn64)
