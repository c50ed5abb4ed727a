"""Parity vs LAPACK and timing of the d <= 64 eigensolver (solver + rotation replay), per solver variant:

    python tools/check_replay_variant.py                           # default: jacobi_d64_kernel at d = 64 + lean replay
    EDRGP_JACOBI_VARIANT=0 python tools/check_replay_variant.py    # the general one-sided kernel at d = 64 too
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from edrgp_b200 import ops
for d in (1, 2, 3, 10, 31, 32, 33, 63, 64):
    rng = np.random.RandomState(d)
    G = rng.standard_normal((3 * d + 5, d)) * np.linspace(3.0, 0.1, d)
    C = G.T.dot(G)
    ev, cp = ops.eigh(torch.as_tensor(C, device='cuda'))
    ev, cp = ev.cpu().numpy(), cp.cpu().numpy()
    lam = np.linalg.eigvalsh(C)[::-1]
    assert np.allclose(ev, lam, rtol=1e-12, atol=1e-12 * lam[0]), d
    assert np.max(np.abs(cp.dot(cp.T) - np.eye(d))) < 1e-12, d
    assert np.max(np.abs(cp.dot(C).dot(cp.T) - np.diag(ev))) < 1e-11 * lam[0], d
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn(100000, 64, dtype=torch.float64, device='cuda', generator=g) * 0.01
G[:, 0] += torch.randn(100000, dtype=torch.float64, device='cuda', generator=g)
C = G.T @ G
for _ in range(3):
    ops.eigh(C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.eigh(C)
e1.record(); e1.synchronize()
print(json.dumps({'jacobi_variant': os.environ.get('EDRGP_JACOBI_VARIANT', 'default'), 'parity': 'ok', 'eigh_d64_ms': e0.elapsed_time(e1) / 20}))
