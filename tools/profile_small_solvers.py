"""Workload for an ncu capture of the latency-bound solvers: edrgp_posv at m = 512 and edrgp_eigh at d = 64 in both
Jacobi variants (run once per variant: EDRGP_JACOBI_VARIANT is read once per process)."""
import sys, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
m, d = 512, 64
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(m, m + 8, dtype=torch.float64, device='cuda', generator=g); A = A @ A.T + m * torch.eye(m, dtype=torch.float64, device='cuda')
b = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
G = torch.randn(100000, d, dtype=torch.float64, device='cuda', generator=g) * 0.01
G[:, 0] += torch.randn(100000, dtype=torch.float64, device='cuda', generator=g)
C = G.T @ G
for _ in range(2):
    ops.posv(A.clone(), b.clone())
    ops.eigh(C)
torch.cuda.synchronize()
