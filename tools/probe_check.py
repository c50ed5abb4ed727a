import sys, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
x = torch.randn(1 << 26, device='cuda'); torch.cuda.synchronize()
for i in range(4):
    print('probe', i, ops.fp64_tensor_peak_tflops(iters=20000, reps=4))
