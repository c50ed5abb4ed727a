"""Numerical study (CPU, NumPy only) of an exact-product int8-sliced SYRK for P = Kfu^T Kfu -- the "one remaining big
lever" of DESIGN.md section 8: the FP64 tensor pipe runs the symmetric reduction at 92 % of its 37 TFLOP/s peak, while
the INT8 tensor cores of a B200 offer 4.5 POP/s.

Scheme.  Kfu entries lie in [0, sf2]; with t = K / (2 sf2) + 1 in [1, 1.5] the 52 mantissa bits of t ARE the fixed-point
fraction of K / (2 sf2), so its bytes are unsigned 8-bit slices q_0 .. q_6 (q_s weighs 2^-8(s+1); the conversion is one
FP64 fma and byte permutes).  Slice pairs with s + s' = g share a weight, their int8 x int8 products accumulate EXACTLY in
32-bit integers over <= 16 K rows, and P = sf2^2 sum_g 2^-8(g+2) A_g with the pairs s + s' >= S dropped.  This script
measures, for S = 4 .. 7 slices, the error that reaches P, the posterior weights and the gradients' Gram matrix
-- against the FP64 statistics -- and counts the int8 GEMMs needed.

    python tools/int8_slice_syrk_study.py            # prints a table (kept in profiles/r02_int8_slice_syrk_study.txt)
"""
import numpy as np


def rbf(X, Z, ell, sf2):
    Xs, Zs = X / ell, Z / ell
    r2 = np.clip((Xs ** 2).sum(1)[:, None] + (Zs ** 2).sum(1)[None, :] - 2 * Xs.dot(Zs.T), 0, None)
    return sf2 * np.exp(-0.5 * r2)


def slices(K, sf2, S):
    t = K / (2.0 * sf2) + 1.0
    mant = t.view(np.uint64) & ((1 << 52) - 1)
    out = []
    for s in range(S):
        shift = 52 - 8 * (s + 1)
        q = (mant >> shift) & 0xFF if shift >= 0 else (mant << -shift) & 0xFF
        out.append(q.astype(np.float64))                  # (products and their sums over n rows stay exact in FP64)
    return out


def sliced_syrk(K, sf2, S):
    q = slices(K, sf2, S)
    P = np.zeros((K.shape[1], K.shape[1]), dtype=np.longdouble)
    gemms = 0.0
    for a in range(S):
        for b in range(a, S):
            if a + b >= S:
                continue
            A = q[a].T.dot(q[b])                          # exact (what the int32 accumulators hold per 16 K rows)
            A = A + A.T if a != b else A
            gemms += 1.0 if a != b else 0.5               # a diagonal pair is itself symmetric: upper triangle only
            P += np.ldexp(A.astype(np.longdouble), -8 * (a + b + 2))
    return (P * 4.0 * sf2 * sf2).astype(np.float64), gemms


def main():
    rng = np.random.RandomState(0)
    n, d, m, sf2, noise = 60000, 64, 256, 1.3, 0.1
    X = rng.standard_normal((n, d))
    B = np.linalg.qr(rng.standard_normal((d, 3)))[0]
    y = np.tanh(X.dot(B)).sum(1) + 0.05 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    Z = X[:m].copy()
    ell = np.sqrt(d) * (1 + 0.5 * rng.uniform(size=d))
    K = rbf(X, Z, ell, sf2)
    Kuu = rbf(Z, Z, ell, sf2) + 1e-8 * np.eye(m)
    beta = 1.0 / noise
    b = K.T.dot(y)

    def downstream(P):
        alpha = np.linalg.solve(Kuu + beta * P, beta * b)
        W = K * alpha
        G = (W.dot(Z) - W.sum(1)[:, None] * X) / ell ** 2
        return alpha, G.T.dot(G)

    P64 = K.T.dot(K)
    Pld = (K.astype(np.longdouble).T.dot(K.astype(np.longdouble))).astype(np.float64)     # reference for FP64's own noise
    a64, C64 = downstream(P64)
    aref, Cref = downstream(Pld)

    def rel(a, r):
        return float(np.max(np.abs(a - r)) / np.max(np.abs(r)))

    print("n=%d d=%d m=%d; errors are max-norm relative to the extended-precision statistics" % (n, d, m))
    print("%-26s %10s %10s %10s %12s" % ("statistics", "P", "alpha", "C = G^T G", "int8 GEMMs"))
    print("%-26s %10.1e %10.1e %10.1e %12s" % ("FP64 (DMMA order-free ref)", rel(P64, Pld), rel(a64, aref), rel(C64, Cref), "-"))
    for S in (4, 5, 6, 7):
        P, gemms = sliced_syrk(K, sf2, S)
        a, C = downstream(P)
        w, _ = np.linalg.eigh(Kuu + beta * P)
        print("%-26s %10.1e %10.1e %10.1e %12.1f   min eig(S) %.2e" % ("int8 slices, S = %d" % S, rel(P, Pld), rel(a, aref),
                                                                     rel(C, Cref), gemms, w[0]))


if __name__ == '__main__':
    main()
