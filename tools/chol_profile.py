"""Phase stamps of chol_step_kernel (tuning build: EDRGP_NVCC_EXTRA=-DCHOL_PROFILE python -m edrgp_b200.build --force)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops, _lib
lib = _lib.load()
m = 512
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(m, m + 8, dtype=torch.float64, device='cuda', generator=g); A = A @ A.T + m * torch.eye(m, dtype=torch.float64, device='cuda')
b = torch.randn(m, dtype=torch.float64, device='cuda', generator=g)
for _ in range(3):
    ops.posv(A.clone(), b.clone())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 8))()
lib.edrgp_debug_chol_prof.argtypes = [ctypes.c_void_p]
print(lib.edrgp_debug_chol_prof(buf))
names = ['stage', 'potf2', 'solve', 'barrier', 'update+store']
prev_end = None
for kb in range(16):
    t = [buf[kb * 8 + i] for i in range(6)]
    gap = (t[0] - prev_end) if prev_end is not None else 0
    print('step %2d  launch gap %6d |' % (kb, gap), '  '.join('%s %6d' % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), '| total', t[5] - t[0])
    prev_end = t[5]
