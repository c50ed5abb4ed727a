"""Device timing of the statistics pass: Kuf block -> HBM, then P, b, yy on the DMMA SYRK (development aid)."""
import sys, json
import torch
sys.path.insert(0, '.')
from edrgp_b200 import ops

n, d, m = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (4_000_000, 64, 512)))
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 524288
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(n, d, dtype=torch.float64, device='cuda', generator=g)
y = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
Z = X[torch.randperm(n, device='cuda', generator=g)[:m]].contiguous()
ell = (d ** 0.5) * (1 + 0.5 * torch.rand(d, dtype=torch.float64, device='cuda', generator=g))
pack = ops.InducingPack(Z, ell)
ldk = m + (m & 1)
Kbuf = torch.empty(min(chunk, n), ldk, dtype=torch.float64, device='cuda')
P = torch.empty(m, m, dtype=torch.float64, device='cuda')
byy = torch.empty(m + 1, dtype=torch.float64, device='cuda')


def ev():
    return torch.cuda.Event(enable_timing=True)


def sweep(tk, ts):
    for i, s in enumerate(range(0, n, chunk)):
        e = min(n, s + chunk)
        a, b, c = ev(), ev(), ev()
        a.record()
        ops.kuf(X[s:e], pack, 1.0, out=Kbuf[:e - s])
        b.record()
        ops.inducing_stats(Kbuf[:e - s], y[s:e], m, P=P, b_yy=byy, accumulate=i > 0)
        c.record()
        tk.append((a, b)); ts.append((b, c))


for _ in range(2):
    sweep([], [])
torch.cuda.synchronize()
best = None
for _ in range(3):
    tk, ts = [], []
    e0, e1 = ev(), ev()
    e0.record(); sweep(tk, ts); e1.record(); e1.synchronize()
    tot = e0.elapsed_time(e1)
    k = sum(a.elapsed_time(b) for a, b in tk); s = sum(a.elapsed_time(b) for a, b in ts)
    if best is None or tot < best[0]:
        best = (tot, k, s)
tot, k, s = best
print(json.dumps({'n': n, 'd': d, 'm': m, 'chunk': chunk, 'ms_total': tot, 'ms_kuf': k, 'ms_syrk': s,
                  'kuf_tflops': n * (2.0 * m * d + m) / k / 1e9, 'syrk_tflops_alg': n * 2.0 * m * m / s / 1e9,
                  'syrk_tflops_exec': n * 1.0 * m * (m + 128) / s / 1e9,
                  'stats_frac_alg_of_37.19': n * (2.0 * m * d + 2.0 * m * m + 3 * m) / tot / 1e9 / 37.19}))
