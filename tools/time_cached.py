"""Device timing of the cached-Kfu gradient kernel vs the recompute kernel (development aid)."""
import sys, json
import torch
sys.path.insert(0, '.')
from edrgp_b200 import ops

n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (1_000_000, 64, 512)))
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].contiguous()
ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
plain = ops.InducingPack(Z, ell)
K = torch.empty(n, m, dtype=torch.float64, device='cuda')
ops.kuf(X, plain, 1.0, out=K)
pack = ops.InducingPack(Z, ell, alpha, 1.0)
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
tc = t(lambda: ops.grad_gram_cached(X, K, pack, 1.0, want_G=False))
tr = t(lambda: ops.grad_gram(X, pack, want_G=False))
print(json.dumps({'n': n, 'ms_cached': tc, 'ms_recompute': tr, 'cached_tflops': n * (2.0 * m * d + 2 * d * d) / tc / 1e9,
                  'cached_GBs': n * (m + d) * 8 / tc / 1e6}))
