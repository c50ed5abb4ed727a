import sys, torch, json
sys.path.insert(0, '.')
from edrgp_b200 import ops
n, m = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device='cuda').manual_seed(0)
K = torch.rand(n, m, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
a = t(lambda: ops.syrk(K)); b = t(lambda: ops.inducing_stats(K, y, m))
print(json.dumps({'n': n, 'm': m, 'ms_syrk': a, 'ms_stats': b, 'exec_tflops_syrk': n * m * (m + 128.0) / a / 1e9}))
