"""Shard-sum parity at FIXED hyper-parameters, run under torchrun (one rank per GPU): SURVEY.md section 4(iv).

Every rank fits its row shard; the all-reduced statistics {P, b, y^T y}, the Gram matrix C of the
posterior-mean gradients, what consumes alpha (posterior mean, gradients) and the EDR directions must equal a
single-process fit over all rows within FP64 summation-order noise, and be bit-identical across ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/check_multigpu_fixed.py [--n 1000003 --d 64 --m 512]

Rank 0 prints one JSON line (kept under profiles/) and asserts: P, b, yy <= 1e-12 relative (max norm); C and
the gradients <= 1e-10 (alpha amplifies the noise of P by the conditioning of Kuu + beta P; its consumers see a
small part of that); principal angle of the leading direction <= 1e-9.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edrgp_b200 as eb                       # noqa: E402
from edrgp_b200 import dist, model            # noqa: E402
from edrgp_b200.utils import principal_angle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=1_000_003)
ap.add_argument('--d', type=int, default=64)
ap.add_argument('--m', type=int, default=512)
ap.add_argument('--chunk-rows', type=int, default=65536)
args = ap.parse_args()

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
n, d, m = args.n, args.d, args.m

rng = np.random.RandomState(0)
X = rng.standard_normal((n, d))
Bm = np.linalg.qr(rng.standard_normal((d, 3)))[0]
y = np.tanh(X.dot(Bm)).sum(1) * 1.7 + 0.4 + 0.05 * rng.standard_normal(n)
Z = X[rng.permutation(n)[:m]].copy()
ell = np.sqrt(d) * (1. + 0.5 * np.random.RandomState(1).uniform(size=d))


def fit(Xr, yr):
    mod = model.SparseGPRegression(Xr, yr[:, None], kernel=model.RBF(d, 1.0, ell, ARD=True), Z=Z, normalizer=True,
                                   noise_var=0.1, chunk_rows=args.chunk_rows)
    G, C = mod.gradient_gram(want_G=True, want_C=True, reduce=True)
    tr = eb.GramEighTransformer().fit_gram(C, n)
    P, byy = mod._stats
    return {'P': P.cpu().numpy(), 'byy': byy.cpu().numpy(), 'C': C.cpu().numpy(), 'alpha': mod.alpha.cpu().numpy(),
            'G': G.cpu().numpy(), 'comps': tr.components_, 'lam': tr.subspace_variance_,
            'mean': mod.normalizer.mean, 'std': mod.normalizer.std, 'll': float(mod.log_likelihood()[0, 0]),
            'mu': mod.predict(X[:2000], want_variance=False)[0][:, 0]}


lo, hi = dist.shard_bounds(n)
for _ in range(3):            # three sweeps: both copies of every exchange payload get used, epochs move on
    mine = fit(X[lo:hi], y[lo:hi])
collectives = 'peer (NVLink exchange buffers)' if dist.peer_exchange(m, d + (d & 1)) is not None else 'nccl (torch.distributed)'


def identical(a):
    t = torch.as_tensor(np.ascontiguousarray(a), device='cuda')
    got = [torch.empty_like(t) for _ in range(world)]
    tdist.all_gather(got, t)
    return all(torch.equal(got[0], g) for g in got)


same = {k: identical(mine[k]) for k in ('P', 'byy', 'C', 'alpha', 'comps', 'mu')}


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


if rank == 0:
    with dist.local_only():
        one = fit(X, y)
    out = {'world': world, 'n': n, 'd': d, 'm': m, 'collectives': collectives, 'identical_across_ranks': same,
           'rel_P': rel(mine['P'], one['P']), 'rel_b': rel(mine['byy'][:m], one['byy'][:m]),
           'rel_yy': abs(mine['byy'][m] - one['byy'][m]) / one['byy'][m],
           'rel_mean': abs(mine['mean'] - one['mean']) / abs(one['mean']), 'rel_std': abs(mine['std'] - one['std']) / one['std'],
           'rel_C': rel(mine['C'], one['C']), 'rel_alpha': rel(mine['alpha'], one['alpha']),
           'rel_G_rank0_rows': rel(mine['G'], one['G'][lo:hi]), 'rel_posterior_mean': rel(mine['mu'], one['mu']),
           'rel_loglik': abs(mine['ll'] - one['ll']) / abs(one['ll']),
           'rel_eigenvalues': float(np.max(np.abs(mine['lam'] - one['lam'])) / one['lam'][0]),
           'angle_leading_direction': principal_angle(mine['comps'][:1], one['comps'][:1])}
    print(json.dumps(out))
    assert all(same.values()), same
    assert out['rel_P'] < 1e-12 and out['rel_b'] < 1e-12 and out['rel_yy'] < 1e-12
    assert out['rel_C'] < 1e-10 and out['rel_G_rank0_rows'] < 1e-10 and out['rel_posterior_mean'] < 1e-10
    assert out['rel_loglik'] < 1e-11
    assert out['angle_leading_direction'] < 1e-9
tdist.barrier()
tdist.destroy_process_group()
