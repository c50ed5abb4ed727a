import sys, torch, json, ctypes
sys.path.insert(0, '.')
from edrgp_b200 import ops, _lib
lib = _lib.load()
d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn(100000, d, dtype=torch.float64, device='cuda', generator=g) * 0.01
G[:, 0] += torch.randn(100000, dtype=torch.float64, device='cuda', generator=g)
for name, C in (('dominant', G.T @ G), ('flat', torch.randn(500, d, dtype=torch.float64, device='cuda', generator=g).T @ torch.randn(500, d, dtype=torch.float64, device='cuda', generator=g))):
    C = 0.5 * (C + C.T) if name == 'flat' else C
    if name == 'flat':
        C = C @ C.T
    A = C.clone(); ev = torch.empty(d, dtype=torch.float64, device='cuda'); cp = torch.empty(d, d, dtype=torch.float64, device='cuda')
    ws = torch.empty(lib.edrgp_eigh_workspace_bytes(d) // 8 + 8, dtype=torch.float64, device="cuda"); sw = torch.zeros(1, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(5):
        A.copy_(C); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.edrgp_eigh(A.data_ptr(), d, ev.data_ptr(), cp.data_ptr(), sw.data_ptr(), ws.data_ptr(), st); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ref = torch.linalg.eigvalsh(C).flip(0)
    print(json.dumps({'case': name, 'd': d, 'ms': best, 'sweeps': int(sw[0]), 'eval_err': float((ev - ref).abs().max() / ref[0])}))
