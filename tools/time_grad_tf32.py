import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
n, d, m = 524288, 64, 512
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].clone()
ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
K, _ = ops.kuf(X, ops.InducingPack(Z, ell), 1.0)
Gb = torch.empty(n, d, dtype=torch.float64, device='cuda')
fn = lambda: ops.grad_tf32(X, K, Z, ell, alpha, 1.0, 1.0, G_out=Gb)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fn()
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({'rows': n, 'ms': ms, 'read_GBs': n * (m + d) * 8 / ms / 1e6}), flush=True)
