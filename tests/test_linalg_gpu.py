"""Parity of the m x m solve chain, SYRK and the Jacobi eigensolver against the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op          # noqa: E402
from oracle import gpy_restatement as gpy  # noqa: E402


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("n,k", [(1000, 20), (5000, 256), (20000, 512), (777, 34), (333, 130), (50000, 128)])
def test_syrk_matches_numpy(n, k):
    from edrgp_b200 import ops
    rng = np.random.RandomState(n + k)
    A = rng.standard_normal((n, k))
    C = ops.syrk(_dev(A)).cpu().numpy()
    ref = A.T.dot(A)
    assert _relerr(C, ref) < 1e-12
    assert np.array_equal(C, C.T)


@pytest.mark.parametrize("n,m", [(1000, 20), (4097, 256), (33333, 512), (777, 33), (15, 130), (1, 6)])
def test_inducing_stats_match_numpy_and_accumulate(n, m):
    """P = K^T K, b = K^T y, yy = y^T y in one deterministic pass; two row chunks accumulate."""
    from edrgp_b200 import ops
    rng = np.random.RandomState(n + m)
    ldk = m + (m & 1)
    K = rng.standard_normal((n, ldk))
    y = rng.standard_normal(n)
    P, byy = ops.inducing_stats(_dev(K), _dev(y), m)
    Kc = K[:, :m]
    assert _relerr(P.cpu().numpy(), Kc.T.dot(Kc)) < 1e-12
    assert _relerr(byy[:m].cpu().numpy(), Kc.T.dot(y)) < 1e-12
    assert abs(float(byy[m]) - y.dot(y)) < 1e-12 * y.dot(y)
    assert np.array_equal(P.cpu().numpy(), P.cpu().numpy().T)
    if n >= 2:
        h = n // 2 + (n // 2) % 2          # even split keeps y 16-byte aligned
        if 0 < h < n:
            P2, byy2 = ops.inducing_stats(_dev(K[:h]), _dev(y[:h]), m)
            ops.inducing_stats(_dev(K[h:]), _dev(y[h:]), m, P=P2, b_yy=byy2, accumulate=True)
            assert _relerr(P2.cpu().numpy(), Kc.T.dot(Kc)) < 1e-12
            assert _relerr(byy2[:m].cpu().numpy(), Kc.T.dot(y)) < 1e-12
    # run-to-run determinism
    P3, byy3 = ops.inducing_stats(_dev(K), _dev(y), m)
    assert torch.equal(P, P3) and torch.equal(byy, byy3)


@pytest.mark.parametrize("n,ka,kb", [(1000, 20, 6), (5000, 256, 64), (20000, 512, 64), (777, 34, 10), (4000, 130, 200)])
def test_gemm_tn_matches_numpy(n, ka, kb):
    from edrgp_b200 import ops
    rng = np.random.RandomState(n + ka)
    A = rng.standard_normal((n, ka))
    B = rng.standard_normal((n, kb))
    C = ops.gemm_tn(_dev(A), _dev(B))
    assert _relerr(C.cpu().numpy(), A.T.dot(B)) < 1e-12
    C = ops.gemm_tn(_dev(A), _dev(B), out=C, accumulate=True)
    assert _relerr(C.cpu().numpy(), 2 * A.T.dot(B)) < 1e-12


@pytest.mark.parametrize("n,d,m", [(500, 10, 20), (2000, 6, 25), (4096, 32, 256), (3000, 64, 512), (900, 7, 33)])
def test_stats_and_solve_match_oracle(n, d, m):
    from edrgp_b200 import ops
    w = op.make_workload(n, d, m, seed=n + m)
    X, y, Z, ell, sf2, noise = w['X'], w['y'], w['Z'], w['ell'], 1.2, w['noise']
    pack = ops.InducingPack(_dev(Z), _dev(ell))
    K, b = ops.kuf(_dev(X), pack, sf2, y=_dev(y))
    Kfull = K if K.is_contiguous() else None
    # P through the SYRK kernel on the stored Kfu (even leading dimension)
    ldk = m + (m & 1)
    Kbuf = torch.zeros(n, ldk, dtype=torch.float64, device='cuda')
    Kbuf[:, :m] = K
    P = ops.syrk(Kbuf, m)
    Pref, bref, yy = op.inducing_stats_chunked(X, y, Z, ell, sf2)
    assert _relerr(P.cpu().numpy(), Pref) < 1e-11
    assert _relerr(b.cpu().numpy(), bref) < 1e-11
    Kmm = ops.kmm(pack, sf2)
    Kmm_ref = op.kuu(Z, ell, sf2)
    assert _relerr(Kmm.cpu().numpy(), Kmm_ref) < 1e-12
    res = ops.solve(Kmm, P, b, 1.0 / noise)
    torch.cuda.synchronize()
    assert res.info.cpu().tolist() == [0, 0]
    ref = op.solve_from_stats(Kmm_ref, Pref, bref, yy, n, sf2, noise)
    # alpha amplifies rounding by cond(Kuu + beta P); compare what the path consumes: mu = Kfu alpha
    mu = K.cpu().numpy().dot(res.alpha.cpu().numpy())
    mu_ref = op.kuf_faithful(X, Z, ell, sf2).dot(ref['alpha'])
    assert _relerr(mu, mu_ref) < 1e-8
    sc = res.scalars.cpu().numpy()
    assert abs(sc[0] - np.trace(ref['A'])) < 1e-9 * abs(np.trace(ref['A']))
    assert abs(sc[1] - np.sum(np.log(np.diag(ref['LB'])))) < 1e-9 * abs(np.sum(np.log(np.diag(ref['LB']))))
    beta = 1.0 / noise
    bound = (-0.5 * n * (np.log(2 * np.pi) - np.log(beta)) - 0.5 * beta * yy - 0.5 * (beta * n * sf2 - sc[0])
             - sc[1] + 0.5 * sc[2])
    assert abs(bound - ref['bound']) < 1e-9 * abs(ref['bound'])
    # Cholesky factors themselves
    Lm = np.tril(res.Lm.cpu().numpy())
    assert _relerr(Lm, ref['Lm']) < 1e-9


@pytest.mark.parametrize("m,nrhs", [(20, 1), (100, 7), (512, 512), (33, 65)])
def test_trsm(m, nrhs):
    from edrgp_b200 import ops
    rng = np.random.RandomState(m)
    A = rng.standard_normal((m, m)); A = A.dot(A.T) + m * np.eye(m)
    L = np.linalg.cholesky(A)
    B = rng.standard_normal((m, nrhs))
    for trans in (False, True):
        X = ops.trsm(_dev(L), _dev(B), trans).cpu().numpy()
        ref = gpy.dtrtrs(L, B, lower=1, trans=int(trans))
        assert _relerr(X, ref) < 1e-11


@pytest.mark.parametrize("m", [1, 5, 32, 33, 64, 100, 257, 512, 1000])
def test_posv_matches_lapack(m):
    """One-launch-per-step Cholesky solve (right-hand side carried as an extra row): solution, factor,
    untouched upper triangle of the factor buffer, run-to-run determinism."""
    from edrgp_b200 import ops
    import scipy.linalg as sl
    rng = np.random.RandomState(m)
    B = rng.standard_normal((m, m + 3))
    A = B.dot(B.T) + 0.1 * m * np.eye(m)
    b = rng.standard_normal(m)
    x, L, info = ops.posv(_dev(A), _dev(b))
    assert int(info.cpu()[0]) == 0
    Lref = np.linalg.cholesky(A)
    assert _relerr(np.tril(L.cpu().numpy()), Lref) < 1e-12
    xref = sl.cho_solve((Lref, True), b)
    assert _relerr(x.cpu().numpy(), xref) < 1e-10
    x2, L2, _ = ops.posv(_dev(A), _dev(b))
    assert torch.equal(x, x2) and torch.equal(torch.tril(L), torch.tril(L2))
    # the same factor as the two-launch-per-step in-place routine
    L3, _ = ops.potrf(_dev(A))
    assert _relerr(np.tril(L.cpu().numpy()), np.tril(L3.cpu().numpy())) < 1e-13


def test_posv_reports_indefinite():
    from edrgp_b200 import ops
    A = np.eye(70); A[40, 40] = -1.0
    _, _, info = ops.posv(_dev(A), _dev(np.ones(70)))
    assert int(info.cpu()[0]) == 41


def test_eigh_general_kernel_matches_lapack():
    """EDRGP_JACOBI_VARIANT=0 selects the general one-sided kernel also at d = 64, where the default is the
    specialised jacobi_d64_kernel (the switch is read once per process: run in a child process)."""
    import os, subprocess, sys
    code = (
        "import numpy as np, torch\n"
        "from edrgp_b200 import ops\n"
        "for d in (3, 10, 33, 64):\n"
        "    rng = np.random.RandomState(d)\n"
        "    G = rng.standard_normal((3 * d + 5, d)) * np.linspace(3.0, 0.1, d)\n"
        "    C = G.T.dot(G)\n"
        "    ev, cp = ops.eigh(torch.as_tensor(C, device='cuda'))\n"
        "    ev, cp = ev.cpu().numpy(), cp.cpu().numpy()\n"
        "    lam = np.linalg.eigvalsh(C)[::-1]\n"
        "    assert np.allclose(ev, lam, rtol=1e-12, atol=1e-12 * lam[0])\n"
        "    assert np.max(np.abs(cp.dot(cp.T) - np.eye(d))) < 1e-12\n"
        "    assert np.max(np.abs(cp.dot(C).dot(cp.T) - np.diag(ev))) < 1e-11 * lam[0]\n"
        "print('ok')\n")
    env = dict(os.environ, EDRGP_JACOBI_VARIANT='0')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("kind", ["flat", "rank_deficient", "one_dominant", "tiny_scale", "huge_scale", "zero"])
def test_eigh_d64_spectra(kind):
    """The default d = 64 path (specialised solver + lean rotation replay) on spectra that stress the sweep count
    and the relative rotation test: flat (clustered), rank 5 of 64, one dominant direction, entries ~1e-200 (the
    solver iterates on a copy scaled by a power of two, or the products of column norms it compares would underflow),
    zero."""
    from edrgp_b200 import ops
    d = 64
    rng = np.random.RandomState(7)
    if kind == "flat":
        G = rng.standard_normal((50000, d))
    elif kind == "rank_deficient":
        G = rng.standard_normal((400, 5)).dot(rng.standard_normal((5, d)))
    elif kind == "one_dominant":
        G = 0.01 * rng.standard_normal((20000, d)); G[:, 0] += rng.standard_normal(20000)
    elif kind == "tiny_scale":
        G = 1e-100 * rng.standard_normal((500, d)) * np.linspace(3.0, 0.1, d)
    elif kind == "huge_scale":
        G = 1e120 * rng.standard_normal((500, d)) * np.linspace(3.0, 0.1, d)
    else:
        G = np.zeros((10, d))
    C = G.T.dot(G)
    evals, comps = ops.eigh(_dev(C))
    evals, comps = evals.cpu().numpy(), comps.cpu().numpy()
    lam = np.linalg.eigvalsh(C)[::-1]
    scale = max(lam[0], 1e-300)
    assert np.all(np.isfinite(evals)) and np.all(np.isfinite(comps))
    assert np.max(np.abs(evals - lam)) <= 1e-12 * scale
    assert np.max(np.abs(comps.dot(comps.T) - np.eye(d))) < 1e-12
    assert np.max(np.abs(comps.dot(C).dot(comps.T) - np.diag(evals))) <= 1e-11 * scale
    assert np.all(np.diff(evals) <= 1e-12 * scale)                       # descending


def test_potrf_reports_indefinite():
    from edrgp_b200 import ops
    A = np.eye(40); A[17, 17] = -1.0
    res = ops.solve(_dev(A), _dev(np.zeros((40, 40))), _dev(np.zeros(40)), 1.0)
    torch.cuda.synchronize()
    assert res.info.cpu().tolist()[0] == 18


@pytest.mark.parametrize("d", [2, 3, 10, 32, 64, 65, 128, 200])
def test_eigh_matches_lapack(d):
    from edrgp_b200 import ops
    rng = np.random.RandomState(d)
    G = rng.standard_normal((3 * d + 5, d)) * np.linspace(3.0, 0.1, d)
    C = G.T.dot(G)
    evals, comps = ops.eigh(_dev(C))
    evals, comps = evals.cpu().numpy(), comps.cpu().numpy()
    lam = np.linalg.eigvalsh(C)[::-1]
    assert np.allclose(evals, lam, rtol=1e-12, atol=1e-12 * lam[0])
    assert np.max(np.abs(comps.dot(comps.T) - np.eye(d))) < 1e-12
    assert np.max(np.abs(comps.dot(C).dot(comps.T) - np.diag(evals))) < 1e-11 * lam[0]
    k = min(3, d)
    cref, _, _ = op.edr_from_gram(C, k)
    assert op.principal_angle(comps[:k], cref) < 1e-6
