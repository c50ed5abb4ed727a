"""Phase stamps of one Jacobi step (tuning build: EDRGP_NVCC_EXTRA=-DJAC_PROFILE python -m edrgp_b200.build --force)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops, _lib
lib = _lib.load()
d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn(5000, d, dtype=torch.float64, device='cuda', generator=g)
C = G.T @ G
for _ in range(2):
    ops.eigh(C)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 4 * 8))()
lib.edrgp_debug_jac_prof.argtypes = [ctypes.c_void_p]
print(lib.edrgp_debug_jac_prof(buf))
names = ['loads+dots', 'shuffles', 'test', 'params', 'rotate+store', 'barrier']
for who in range(2):
    for st in range(4):
        t = [buf[(who * 4 + st) * 8 + i] for i in range(6)]
        nxt = buf[(who * 4 + st + 1) * 8] if st < 3 else None
        print('thread %d step %d:' % (who, st + 4), '  '.join('%s %5d' % (n, t[i + 1] - t[i]) for i, n in enumerate(names[:5]) if True), '| step total', (nxt - t[0]) if nxt else t[5] - t[0])
