"""Validation of the lean rotation-replay kernel (EDRGP_JACOBI_REPLAY=1; written at the end of round 1 after the GPU
budget was spent, so it is OFF by default and has not run on a GPU yet).

    EDRGP_JACOBI_REPLAY=1 python tools/check_replay_variant.py     # parity vs LAPACK + timing
    python tools/check_replay_variant.py                           # the default replay kernel, for comparison
    EDRGP_JACOBI_VARIANT=5 python tools/check_replay_variant.py    # the d = 64 specialised solver (also unvalidated)
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
print(json.dumps({'replay_variant': os.environ.get('EDRGP_JACOBI_REPLAY', '0'), 'jacobi_variant': os.environ.get('EDRGP_JACOBI_VARIANT', '0'), 'parity': 'ok', 'eigh_d64_ms': e0.elapsed_time(e1) / 20}))
