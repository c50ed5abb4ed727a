"""Multi-rank behaviour of what sits behind the fit -- ``refit`` (Gram form, gathered rows, seeded subsample),
``first_gradients_rows``, ``BlockEDR`` -- run under torchrun (one rank per GPU); rank 0 compares with a single-process
fit over all rows (reference semantics: edrgp/base.py:202-239,520-766; edrgp/edr.py:115-140).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/check_multigpu_refit.py
"""
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edrgp_b200 as eb                                   # noqa: E402
from edrgp_b200 import dist, model                        # noqa: E402
from edrgp_b200.utils import principal_angle              # noqa: E402


class EconomySVDTransformer(object):
    """Any host transformer with fit + components_ (here: uncentred PCA by economy SVD, sklearn-clonable)."""

    def __init__(self, n_components=None):
        self.n_components = n_components

    def get_params(self, deep=True):
        return {'n_components': self.n_components}

    def set_params(self, **params):
        for key, v in params.items():
            setattr(self, key, v)
        return self

    def fit(self, G, y=None):
        _, _, Vh = np.linalg.svd(np.asarray(G, dtype=np.float64), full_matrices=False)
        self.components_ = Vh[:self.n_components or Vh.shape[0]]
        return self


rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
tdist.init_process_group('nccl', device_id=torch.device('cuda', local))

n, d, m, k = 60_001, 8, 48, 2
rng = np.random.RandomState(0)
X = rng.standard_normal((n, d)) * np.linspace(1.5, 0.6, d)
B = np.linalg.qr(rng.standard_normal((d, k)))[0]
y = np.tanh(X.dot(B)).dot([1.0, 0.6]) + 0.05 * rng.standard_normal(n)
Z = X[rng.permutation(n)[:m]].copy()
ell = np.sqrt(d) * (1. + 0.5 * np.random.RandomState(1).uniform(size=d))


def est():
    # (the estimator is refitted on the projected rows: kernel and inducing inputs follow the width it is given)
    return eb.SparseGaussianProcessRegressor('RBF', {'ARD': True}, num_inducing=m, method='fixed', noise_var=0.1,
                                             chunk_rows=8192)


def run(Xr, yr):
    out = {}
    np.random.seed(5)                     # same seed on every rank: same global draw of the inducing rows
    edr = eb.EffectiveDimensionalityReduction(est(), eb.GramEighTransformer(), n_components=k, normalize=True).fit(Xr, yr)
    out['components'] = edr.components_
    edr.refit(eb.GramEighTransformer(n_components=k))
    out['refit_gram'] = (edr.refit_components_, edr.refit_subspace_variance_ratio_)
    edr.refit(EconomySVDTransformer(n_components=k))                                   # all rows gathered
    out['refit_host'] = (edr.refit_components_, edr.refit_subspace_variance_ratio_)
    rows = np.arange(0, n, 7)
    edr.refit(EconomySVDTransformer(n_components=k), rows)
    out['refit_rows'] = (edr.refit_components_, edr.refit_subspace_variance_ratio_)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        edr.refit(EconomySVDTransformer(n_components=k), max_rows=5000)                # seeded subsample, announced
        out['subsample_warned'] = any('subsample' in str(x.message) for x in w)
    out['refit_sub'] = (edr.refit_components_, edr.refit_subspace_variance_ratio_, edr.refit_rows_)
    G, used = edr.first_gradients_rows(rows)
    out['rows_gathered'] = G
    blk = eb.BlockEDR(est(), eb.GramEighTransformer(), n_components=[1, 2], blocks=[[0, 1, 2, 3], [4, 5, 6, 7]]).fit(Xr, yr)
    out['block'] = (blk.components_, blk.subspace_variance_ratio_)
    return out


lo, hi = dist.shard_bounds(n)
mine = run(X[lo:hi], y[lo:hi])


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def signed(a, b):
    s = np.sign(np.sum(a * b, axis=1))
    return rel(a * s[:, None], b)


if rank == 0:
    with dist.local_only():
        one = run(X, y)
    res = {'world': world,
           'angle_components': principal_angle(mine['components'], one['components']),
           'subsample_warned': mine['subsample_warned'],
           'subsample_rows_identical': bool(np.array_equal(mine['refit_sub'][2], one['refit_sub'][2])),
           'rows_gathered': rel(mine['rows_gathered'], one['rows_gathered'])}
    for key in ('refit_gram', 'refit_host', 'refit_rows', 'refit_sub', 'block'):
        res[key + '_components'] = signed(mine[key][0], one[key][0])
        res[key + '_ratio'] = rel(mine[key][1], one[key][1])
    print(json.dumps(res))
    assert res['subsample_warned'] and res['subsample_rows_identical']
    assert res['angle_components'] < 1e-8 and res['rows_gathered'] < 1e-9
    assert all(v < 1e-7 for kk, v in res.items() if kk.endswith('_components') or kk.endswith('_ratio')), res
tdist.barrier()
tdist.destroy_process_group()
