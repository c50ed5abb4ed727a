"""Parity and timing of the INT8 symmetric reduction (edrgp_inducing_stats_i8) against the FP64 DMMA reduction."""
import json
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from edrgp_b200 import ops

import os
res = []
CASES = () if os.environ.get('I8_BIG_ONLY') else ((1000, 8, 64, 1.0), (20000, 16, 300, 1.7), (70001, 64, 512, 0.4))
for (n, d, m, sf2) in CASES:
    g = torch.Generator(device='cuda').manual_seed(n)
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
    Z = X[:m].contiguous()
    ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
    K, _ = ops.kuf(X, ops.InducingPack(Z, ell), sf2)
    K = K.contiguous() if K.shape[1] == m + (m & 1) else K
    Kfull = torch.empty(n, m + (m & 1), dtype=torch.float64, device='cuda')
    ops.kuf(X, ops.InducingPack(Z, ell), sf2, out=Kfull)
    P0, b0 = ops.inducing_stats(Kfull, y, m)
    P1, b1 = ops.inducing_stats_i8(Kfull, y, sf2, m)
    torch.cuda.synchronize()
    Kh = Kfull[:, :m].cpu().numpy().astype(np.longdouble)
    Pref = (Kh.T @ Kh).astype(np.float64) if n <= 20000 else None
    r = {'n': n, 'm': m, 'rel_P_i8_vs_fp64': float((P1 - P0).abs().max() / P0.abs().max()),
         'rel_b': float((b1[:m] - b0[:m]).abs().max() / b0[:m].abs().max()), 'rel_yy': float(abs(b1[m] - b0[m]) / b0[m]),
         'sym': float((P1 - P1.T).abs().max())}
    if Pref is not None:
        r['rel_P_i8_vs_ld'] = float(np.max(np.abs(P1.cpu().numpy() - Pref)) / np.max(np.abs(Pref)))
        r['rel_P_fp64_vs_ld'] = float(np.max(np.abs(P0.cpu().numpy() - Pref)) / np.max(np.abs(Pref)))
    # accumulate mode
    P2, b2 = ops.inducing_stats_i8(Kfull, y, sf2, m, P=P1.clone(), b_yy=b1.clone(), accumulate=True)
    r['accumulate'] = float((P2 - 2 * P1).abs().max() / P1.abs().max())
    res.append(r)
    print(json.dumps(r), flush=True)

n, d, m = 524288, 64, 512
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].contiguous()
ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
K = torch.empty(n, m, dtype=torch.float64, device='cuda')
ops.kuf(X, ops.InducingPack(Z, ell), 1.0, out=K)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


t64 = timed(lambda: ops.inducing_stats(K, y, m))
t8 = timed(lambda: ops.inducing_stats_i8(K, y, 1.0, m))
P0, _ = ops.inducing_stats(K, y, m)
P1, _ = ops.inducing_stats_i8(K, y, 1.0, m)
# extended-precision reference for a 24-column corner (the full product is too slow on the host)
Kc = K[:, :24].cpu().numpy().astype(np.longdouble)
Pld = (Kc.T @ Kc).astype(np.float64)
pmax = float(P0.abs().max())
print(json.dumps({'rows': n, 'fp64_dmma_ms': t64, 'int8_ms': t8, 'rel_P': float((P1 - P0).abs().max() / pmax),
                  'corner24_rel_err_fp64_dmma': float(np.max(np.abs(P0[:24, :24].cpu().numpy() - Pld)) / pmax),
                  'corner24_rel_err_int8x6': float(np.max(np.abs(P1[:24, :24].cpu().numpy() - Pld)) / pmax)}))
