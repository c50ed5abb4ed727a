"""TF32-split gradient contraction vs the FP64 cached kernel: relative error per shape, then timing of 524288 rows."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
shapes = [(128, 64, 128), (1000, 64, 512), (300, 10, 20), (5000, 32, 256), (777, 48, 130)]
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    shapes = shapes[:1]
for n, d, m in shapes:
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    Z = X[:m].clone() if m <= n else torch.randn(m, d, dtype=torch.float64, device='cuda', generator=g)
    ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
    alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
    sf2 = 1.7
    K, _ = ops.kuf(X, ops.InducingPack(Z, ell), sf2)
    K = K.contiguous()
    Gref, _ = ops.grad_gram_cached(X, K, ops.InducingPack(Z, ell, alpha, 0.9, block=64), sf2, want_G=True, want_C=False)
    G = ops.grad_tf32(X, K, Z, ell, alpha, 0.9, sf2)
    torch.cuda.synchronize()
    rel = ((G - Gref).abs().max() / Gref.abs().max()).item()
    print(json.dumps({'n': n, 'd': d, 'm': m, 'rel_err_maxnorm': rel, 'finite': bool(torch.isfinite(G).all())}), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    sys.exit(0)
n, d, m = 524288, 64, 512
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].clone()
ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
K, _ = ops.kuf(X, ops.InducingPack(Z, ell), 1.0)
Gb = torch.empty(n, d, dtype=torch.float64, device='cuda')
p64 = ops.InducingPack(Z, ell, alpha, 1.0, block=64)
for name, fn in (('tf32x3', lambda: ops.grad_tf32(X, K, Z, ell, alpha, 1.0, 1.0, G_out=Gb)),
                 ('tf32x3+syrk', lambda: ops.syrk(ops.grad_tf32(X, K, Z, ell, alpha, 1.0, 1.0, G_out=Gb))),
                 ('fp64 fused G+C', lambda: ops.grad_gram_cached(X, K, p64, 1.0, want_G=False, want_C=True))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({'kernel': name, 'rows': n, 'ms': ms, 'read_GBs': n * (m + d) * 8 / ms / 1e6}), flush=True)
