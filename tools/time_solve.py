import sys, torch, json
sys.path.insert(0, '.')
from edrgp_b200 import ops
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(m, m + 8, dtype=torch.float64, device='cuda', generator=g); A = A @ A.T + m * torch.eye(m, dtype=torch.float64, device='cuda')
b = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
S = torch.empty_like(A); v = torch.empty_like(b)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
cp = t(lambda: S.copy_(A))
print(json.dumps({'m': m, 'potrf_ms': t(lambda: (S.copy_(A), ops.potrf(S))) - cp}))
ops.potrf(S.copy_(A))
print(json.dumps({'m': m, 'trsv_fwd_ms': t(lambda: ops.trsm(S, v.copy_(b), False)), 'trsv_bwd_ms': t(lambda: ops.trsm(S, v.copy_(b), True))}))
print(json.dumps({'m': m, 'posv_ms': t(lambda: ops.posv(S.copy_(A), v.copy_(b))) - cp}))
xp, _, _ = ops.posv(S.copy_(A), v.copy_(b))
print('posv residual', float((A @ xp - b).abs().max()))
ops.potrf(S.copy_(A))
x = b.clone(); ops.trsm(S, x, False); ops.trsm(S, x, True)
print('residual', float((A @ x - b).abs().max()))
