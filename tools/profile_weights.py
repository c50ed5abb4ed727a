import sys, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
n, m = 262144, 512
g = torch.Generator(device='cuda').manual_seed(0)
K = torch.rand(n, m, dtype=torch.float64, device='cuda', generator=g)
M = torch.randn(m, m, dtype=torch.float64, device='cuda', generator=g); M = M + M.T
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
T = torch.empty_like(K); cs = torch.zeros(m, dtype=torch.float64, device='cuda')
for _ in range(3):
    rs = ops.weights(K, M, m, y=y, alpha=alpha, c_ya=2.0, c_km=2.0, T=T, want_rowsum=True, colsum=cs, accumulate=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.weights(K, M, m, y=y, alpha=alpha, c_ya=2.0, c_km=2.0, T=T, want_rowsum=True, colsum=cs, accumulate=True); e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
print('weights ms', ms, 'TF', 2.0 * n * m * m / ms / 1e9)
