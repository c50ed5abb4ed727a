"""CUDA-event times of the n-scale FP64 kernels on one row block of the headline shape (development aid):

    python tools/time_kernels.py [n d m]       -> one JSON line {kernel: ms, ...} + TFLOP/s figures
"""
import json
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
K = torch.empty(n, m + (m & 1), dtype=torch.float64, device='cuda')


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


out = {'n': n, 'd': d, 'm': m}
out['kuf'] = timed(lambda: ops.kuf(X, pack, 1.3, out=K))
out['inducing_stats'] = timed(lambda: ops.inducing_stats(K, y, m))
if d <= 64:
    out['grad_gram_cached'] = timed(lambda: ops.grad_gram_cached(X, K, cpack, 1.3, want_G=False))
if d <= 128:
    out['grad_gram_fused'] = timed(lambda: ops.grad_gram(X, cpack, want_G=False, want_C=d <= 64))
peak = ops.fp64_tensor_peak_tflops()
out['peak_tflops'] = peak
out['kuf_tflops_dmma_only'] = 2.0 * m * d * n / (out['kuf'] * 1e-3) / 1e12
out['kuf_frac_of_peak_dmma_only'] = out['kuf_tflops_dmma_only'] / peak
if 'grad_gram_fused' in out:
    out['fused_tflops'] = (4.0 * m * d + 2.0 * d * d + m) * n / (out['grad_gram_fused'] * 1e-3) / 1e12
    out['fused_frac'] = out['fused_tflops'] / peak
print(json.dumps(out))
