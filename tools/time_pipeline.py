"""Quick device timing of the fused pipeline kernel (development aid; bench.py is the contract)."""
import sys, json
import torch
sys.path.insert(0, '.')
from edrgp_b200 import ops

n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (4_000_000, 64, 512)))
want_G = len(sys.argv) > 4 and sys.argv[4] == 'G'
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[torch.randperm(n, device='cuda', generator=g)[:m]].contiguous()
ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
pack = ops.InducingPack(Z, ell, alpha, 1.0)
Gbuf = torch.empty(n, d, dtype=torch.float64, device='cuda') if want_G else None
for _ in range(2):
    ops.grad_gram(X, pack, want_G=want_G, G_out=Gbuf)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.grad_gram(X, pack, want_G=want_G, G_out=Gbuf); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
flop = n * (4.0 * m * d + 2.0 * d * d + m)
print(json.dumps({'n': n, 'd': d, 'm': m, 'G': want_G, 'ms': ms, 'ms_all': ts, 'pts_per_s': n / ms * 1e3,
                  'tflops_alg': flop / ms / 1e9, 'frac_of_37.19': flop / ms / 1e9 / 37.19}))
