import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
n, d, m = 524288, 64, 512
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].clone()
ell = torch.as_tensor(np.sqrt(d) * (1 + 0.5 * np.random.RandomState(1).uniform(size=d)), device='cuda')
K = torch.empty(n, m, dtype=torch.float64, device='cuda'); ops.kuf(X, ops.InducingPack(Z, ell), 1.0, out=K)
A = torch.randn(m, m, dtype=torch.float64, device='cuda', generator=g); M = 0.5 * (A + A.T) / m
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g); alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
T = torch.empty(n, m, dtype=torch.float64, device='cuda')
fn = lambda: ops.weights_tf32(K, M, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=T, want_rowsum=True)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); e1.synchronize()
print(json.dumps({'rows': n, 'ms': e0.elapsed_time(e1) / 5}))
