"""Value checks at the HEADLINE row count (BASELINE config 3: n = 4M, d = 64, m = 512), where the oracle
cannot run over all rows: what the oracle can do in seconds is one 65 536-row stripe; the rest is pinned
by size-independent identities of the path (the statistics and the Gram matrix are sums over row blocks).

    * stripe: Kfu -> P, b on rows [s, s + 65536) of the 4M-row matrix, and the posterior-mean gradients of
      those rows from the Kfu blocks the FULL fit stored, against oracle/pipeline.py on the same rows;
    * block sums: P, b, y^T y of the full fit (8 blocks of 524 288 rows, accumulated on the device through the
      split-K partials) equal the sum of the eight blocks' own statistics; C = G^T G of the full gradient pass
      equals the sum over the blocks of G_b^T G_b formed from the gradients the same pass wrote;
    * the directions: eigh of the device Gram matrix against LAPACK on the same matrix.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op                  # noqa: E402

N, D, M = 4_000_000, 64, 512
BLOCK = 524288
STRIPE0, STRIPE = 1_700_003, 65536                 # crosses the boundary of blocks 3 and 4 (odd start)


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def test_headline_size_stripe_and_block_sums():
    from edrgp_b200 import model, ops
    import edrgp_b200 as eb
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~25 GB of device memory")
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(77)
    X = torch.randn(N, D, dtype=torch.float64, device=dev, generator=g)
    B = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((D, 3)))[0], device=dev)
    y = torch.tanh(X @ B).sum(1) + 0.05 * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
    Z = X[:M].cpu().numpy()
    ell = np.sqrt(D) * (1. + 0.5 * np.random.RandomState(1).uniform(size=D))
    sf2, noise = 1.0, 0.1

    mod = model.SparseGPRegression(X, y, kernel=model.RBF(D, sf2, ell, ARD=True), Z=Z, normalizer=True,
                                   noise_var=noise, chunk_rows=BLOCK)
    assert mod._Kcache is not None                   # the 16.4 GB Kfu cache path is the one under test
    P_full, byy_full = (t.cpu().numpy() for t in mod._stats)
    mean, std = mod.normalizer.mean, mod.normalizer.std
    yn = mod.Y_normalized

    # ---- block sums of the statistics (fixed order, long double on the host)
    pack = ops.InducingPack(mod._Z_dev, mod._ell_dev)
    Psum = np.zeros((M, M), dtype=np.longdouble)
    bsum = np.zeros(M + 1, dtype=np.longdouble)
    for s in range(0, N, BLOCK):
        e = min(N, s + BLOCK)
        K, _ = ops.kuf(X[s:e], pack, sf2)
        assert torch.equal(K, mod._Kcache[s:e, :M])                     # the stored blocks are what a fresh call writes
        Pb, bb = ops.inducing_stats(K, yn[s:e], M)
        Psum += Pb.cpu().numpy()
        bsum += bb.cpu().numpy()
        del K, Pb, bb
    assert _rel(P_full, Psum.astype(np.float64)) < 1e-12
    assert _rel(byy_full, bsum.astype(np.float64)) < 1e-12
    assert abs(byy_full[M] - N) < 1e-9 * N                              # standardised targets: y^T y = n

    # ---- the stripe against the oracle
    s0, s1 = STRIPE0, STRIPE0 + STRIPE
    Xs, ys = X[s0:s1].cpu().numpy(), yn[s0:s1].cpu().numpy()
    P_ref, b_ref, yy_ref = op.inducing_stats_chunked(Xs, ys, Z, ell, sf2)
    Kd = mod._Kcache[s0:s1]
    if Kd.data_ptr() % 16:                                              # odd start row: the ABI wants 16-byte alignment
        Kd = Kd.clone()
    Pd, bd = ops.inducing_stats(Kd, yn[s0:s1].clone(), M)
    assert _rel(Pd.cpu().numpy(), P_ref) < 1e-12
    assert _rel(bd.cpu().numpy()[:M], b_ref) < 1e-11
    assert abs(float(bd[M]) - yy_ref) < 1e-12 * yy_ref
    K_ref = op.kuf_faithful(Xs[:2048], Z, ell, sf2)
    assert _rel(mod._Kcache[s0:s0 + 2048, :M].cpu().numpy(), K_ref) < 1e-12

    G, C = mod.gradient_gram(want_G=True, want_C=True)
    alpha = mod.alpha.cpu().numpy()
    G_ref = op.gradients_chunked(Xs, Z, ell, sf2, alpha, scale=std)
    assert _rel(G[s0:s1].cpu().numpy(), G_ref) < 1e-10

    # ---- block sums of the Gram matrix, formed independently from the gradients the pass wrote
    Csum = np.zeros((D, D), dtype=np.longdouble)
    for s in range(0, N, BLOCK):
        Gb = G[s:min(N, s + BLOCK)]
        Csum += ops.syrk(Gb).cpu().numpy()
    Cd = C.cpu().numpy()
    assert _rel(Cd, Csum.astype(np.float64)) < 1e-12
    Gh = G[s0:s1].cpu().numpy()
    assert _rel(ops.syrk(G[s0:s1].clone()).cpu().numpy(), Gh.T.dot(Gh)) < 1e-12

    # ---- alpha through what consumes it: the posterior mean of the stripe against a host solve of the SAME
    # reduced statistics (LAPACK Cholesky chain of the oracle)
    sol = op.solve_from_stats(op.kuu(Z, ell, sf2), P_full, byy_full[:M], float(byy_full[M]), N, sf2, noise)
    Kh = mod._Kcache[s0:s0 + 4096, :M].cpu().numpy()
    assert _rel(Kh.dot(alpha), Kh.dot(sol['alpha'])) < 1e-8
    ll = float(mod.log_likelihood()[0, 0])
    assert abs(ll - sol['bound']) < 1e-9 * abs(sol['bound'])

    # ---- the directions
    tr = eb.GramEighTransformer().fit_gram(C, N)
    lam = np.linalg.eigvalsh(Cd)[::-1]
    assert np.allclose(tr.subspace_variance_, lam, rtol=1e-10, atol=1e-12 * lam[0])
    cref, _, _ = op.edr_from_gram(Cd, 1)
    assert op.principal_angle(tr.components_[:1], cref) < 1e-9
