"""TF32-split cross-covariance vs the FP64 kernel: max relative error per shape, then timing of a 524288-row block."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
shapes = [(128, 64, 128), (1000, 64, 512), (300, 10, 20), (5000, 32, 256), (777, 48, 130), (4096, 64, 1024)]
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    shapes = shapes[:1]
for n, d, m in shapes:
    X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
    Z = X[torch.randperm(n, device='cuda', generator=g)[:m] % n].clone() if m <= n else torch.randn(m, d, dtype=torch.float64, device='cuda', generator=g)
    Z = Z + 0.0
    ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
    sf2 = 1.7
    Kref, _ = ops.kuf(X, ops.InducingPack(Z, ell), sf2)
    K = ops.kuf_tf32(X, ops.InducingPackTF32(Z, ell), sf2)
    torch.cuda.synchronize()
    rel = ((K - Kref).abs() / Kref.abs().clamp_min(1e-300)).max().item()
    print(json.dumps({'n': n, 'd': d, 'm': m, 'max_rel_err': rel, 'max_abs_err': (K - Kref).abs().max().item(),
                      'kmin': Kref.min().item(), 'finite': bool(torch.isfinite(K).all())}), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == 'first':
    sys.exit(0)
n, d, m = 524288, 64, 512
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].clone()
ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
K = torch.empty(n, m, dtype=torch.float64, device='cuda')
p32, p64 = ops.InducingPackTF32(Z, ell), ops.InducingPack(Z, ell)
for name, fn in (('tf32x3', lambda: ops.kuf_tf32(X, p32, 1.0, out=K)), ('fp64', lambda: ops.kuf(X, p64, 1.0, out=K))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({'kernel': name, 'rows': n, 'ms': ms, 'write_GBs': n * m * 8 / ms / 1e6}), flush=True)
