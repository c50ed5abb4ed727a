"""Multi-GPU parity check, run under torchrun (one rank per GPU): the n-sharded EDR fit must give
every rank the same components, equal to a single-process fit over all rows within FP64
summation-order noise.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edrgp_b200 as eb                      # noqa: E402
from edrgp_b200 import dist                  # noqa: E402
from edrgp_b200.utils import principal_angle  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
tdist.init_process_group('nccl', device_id=torch.device('cuda', local))

n, d, m, k = 200_003, 12, 64, 2
rng = np.random.RandomState(0)
X = rng.standard_normal((n, d)) * np.linspace(2.0, 0.5, d) + 0.3
B = np.linalg.qr(rng.standard_normal((d, k)))[0]
y = np.tanh(X.dot(B)).sum(1) + 0.05 * rng.standard_normal(n)


def make():
    return eb.EffectiveDimensionalityReduction(
        eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=m, chunk_rows=32768),
        eb.GramEighTransformer(), n_components=k, step=5, normalize=True)


lo, hi = dist.shard_bounds(n)
np.random.seed(11)                            # same seed on every rank: same global Z draw
edr = make().fit(X[lo:hi], y[lo:hi], max_iters=15)
comps = torch.as_tensor(edr.components_, device='cuda')
gathered = [torch.empty_like(comps) for _ in range(world)]
tdist.all_gather(gathered, comps)
same = all(torch.equal(gathered[0], g) for g in gathered)
ll = float(edr.estimator_.estimator_.log_likelihood()[0, 0])

out = {'world': world, 'components_identical_across_ranks': bool(same), 'num_iter': edr.num_iter, 'loglik': ll}
if rank == 0:
    with dist.local_only():
        np.random.seed(11)
        ref = make().fit(X, y, max_iters=15)
    out['angle_vs_single_process'] = principal_angle(edr.components_, ref.components_)
    out['max_abs_component_diff'] = float(np.max(np.abs(np.abs(edr.components_) - np.abs(ref.components_))))
    out['ratio_diff'] = float(np.max(np.abs(edr.subspace_variance_ratio_ - ref.subspace_variance_ratio_)))
    out['loglik_single'] = float(ref.estimator_.estimator_.log_likelihood()[0, 0])
    out['angle_to_truth'] = principal_angle(edr.components_, B.T)
    print(json.dumps(out))
    assert same
    assert out['angle_vs_single_process'] < 1e-6
    assert abs(out["loglik"] - out["loglik_single"]) < 1e-3 * abs(out["loglik_single"])   # L-BFGS trajectories amplify summation-order noise
tdist.barrier()
tdist.destroy_process_group()
