"""Runs the n-scale kernels of the sweep once each after a warm-up, at one 524288-row block of the
headline shape (d=64, m=512), for ncu captures: kuf_kernel (pipelined), gemm_tn_kernel (symmetric),
grad_gram_kernel (cached-Kfu variant and the fused recompute variant), the TF32-split pair
(kuf_tf32_kernel, grad_tf32_kernel) and the d x d Jacobi eigensolver."""
import sys
import torch
sys.path.insert(0, '.')
from edrgp_b200 import ops

n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (524288, 64, 512)))
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[:m].contiguous()
ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
alpha = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
pack = ops.InducingPack(Z, ell)
cpack = ops.InducingPack(Z, ell, alpha, 1.0, block=64)
K = torch.empty(n, m, dtype=torch.float64, device='cuda')
p32 = ops.InducingPackTF32(Z, ell)
for it in range(2):
    ops.kuf_tf32(X, p32, 1.0, out=K)
    G32 = ops.grad_tf32(X, K, Z, ell, alpha, 1.0, 1.0)
    Mw = torch.randn(m, m, dtype=torch.float64, device='cuda', generator=g); Mw = 0.5 * (Mw + Mw.T) / m
    Tw = torch.empty(n, m, dtype=torch.float64, device='cuda')
    ops.weights_tf32(K, Mw, m, y=y, alpha=alpha, c_ya=0.7, c_km=2.0, T=Tw, want_rowsum=True)
    del Tw
    ops.kuf(X, pack, 1.0, out=K)
    P, byy = ops.inducing_stats(K, y, m)
    G, C = ops.grad_gram_cached(X, K, cpack, 1.0, want_G=False)
    G2, C2 = ops.grad_gram(X, cpack, want_G=False)
    ev, comps = ops.eigh(C)
torch.cuda.synchronize()
print("ok", float(P[0, 0]), float(C[0, 0]), float(C2[0, 0]))
