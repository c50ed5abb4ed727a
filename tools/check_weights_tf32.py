"""TF32-split weights kernel vs the FP64 one: relative error per shape, then timing of a 524288-row block."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
shapes = [(128, 16, 128), (1000, 64, 512), (300, 10, 20), (5000, 32, 256), (777, 48, 130)]
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    shapes = shapes[:1]
for n, d, m in shapes:
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    Z = torch.randn(m, d, dtype=torch.float64, device='cuda', generator=g)
    ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
    ldk = m + (m & 1)
    K = torch.zeros(n, ldk, dtype=torch.float64, device='cuda')
    ops.kuf(X, ops.InducingPack(Z, ell), 1.3, out=K)
    A = torch.randn(m, m, dtype=torch.float64, device='cuda', generator=g); M = ops.even_ld(0.5 * (A + A.T) / m)
    y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
    alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
    Tref = torch.zeros(n, ldk, dtype=torch.float64, device='cuda'); T = torch.zeros(n, ldk, dtype=torch.float64, device='cuda')
    rs_ref = ops.weights(K, M, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=Tref, want_rowsum=True)
    rs = ops.weights_tf32(K, M, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=T, want_rowsum=True)
    torch.cuda.synchronize()
    print(json.dumps({'n': n, 'm': m, 'T_rel_err_maxnorm': ((T - Tref).abs().max() / Tref.abs().max()).item(),
                      'rowsum_rel_err': ((rs - rs_ref).abs().max() / rs_ref.abs().max()).item(),
                      'finite': bool(torch.isfinite(T).all())}), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    sys.exit(0)
n, d, m = 524288, 64, 512
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].clone()
ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
K = torch.empty(n, m, dtype=torch.float64, device='cuda'); ops.kuf(X, ops.InducingPack(Z, ell), 1.0, out=K)
A = torch.randn(m, m, dtype=torch.float64, device='cuda', generator=g); M = 0.5 * (A + A.T) / m
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g); alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
T = torch.empty(n, m, dtype=torch.float64, device='cuda')
for name, fn in (('tf32x3', lambda: ops.weights_tf32(K, M, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=T, want_rowsum=True)),
                 ('fp64', lambda: ops.weights(K, M, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=T, want_rowsum=True))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({'kernel': name, 'rows': n, 'ms': ms, 'tflops_alg': 2.0 * n * m * m / ms / 1e9}), flush=True)
