"""Device-resident sparse GP regression model: the object edr-gp keeps as ``estimator_``.

Stands in for ``GPy.models.SparseGPRegression`` as edr-gp builds and uses it
(``edrgp/gp_model/regression.py:153-157``: construction; ``edrgp/gp_model/base.py:65-69``:
``optimize`` / ``optimize_restarts``; ``:187,206``: ``predict``; ``:222``: ``predictive_gradients``;
``edrgp/tests/test_edr.py:49-50``: ``log_likelihood()[0][0]``).  The arithmetic is GPy's VarDTC
(VFE) inference for an RBF kernel with a Gaussian likelihood, regrouped so that everything that
scales with the number of points n runs as row-sharded CUDA kernels through the C ABI
(``include/edrgp_b200.h``) and only m x m / d x d quantities are reduced across GPUs:

    pass 1  Kfu block -> HBM; P = Kfu^T Kfu, b = Kfu^T y, y^T y           [edrgp_kuf, edrgp_inducing_stats]
            all-reduce {P, b, yy}; Kuu; Cholesky chain -> alpha, bound   [edrgp_kmm, edrgp_solve]
    pass 2  (optimisation only) T = Kfu o dL/dKfu; T^T X, T 1, T^T 1, sum_i (T 1)_i x_i^2
                                                       [edrgp_weights, edrgp_gemm_tn, edrgp_col_moments]
            all-reduce {T^T X, T^T 1, moments}; host assembles dL/d{Z, variance, lengthscale, noise}
    sweep   posterior-mean gradients and their Gram matrix                [edrgp_grad_gram]

There is no CPU fallback: every method needs the CUDA library and a CUDA device.
"""
import weakref

import numpy as np
import torch
from scipy import optimize as sopt

from . import dist, ops

F64 = torch.float64
CONST_JITTER = 1e-8                      # GPy VarDTC.const_jitter
_LIM_VAL = 36.0                          # paramz.transformations._lim_val
_LOG_LIM_VAL = float(np.log(np.finfo(np.float64).max))


# ------------------------------------------------------------------------------------------------
# paramz Logexp transform (positive parameters are optimised through log(1 + exp(x)))
# ------------------------------------------------------------------------------------------------
def _logexp_f(x):
    return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))


def _logexp_finv(f):
    return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))


def _logexp_gradfactor(f, df):
    return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))


class RBF(object):
    """Hyper-parameter holder with the surface of ``GPy.kern.RBF`` that edr-gp touches
    (``edrgp/gp_model/base.py:111-147`` builds it from ``'RBF'`` + ``kernel_options``)."""

    name = 'rbf'

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, **unused):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.variance = float(variance)
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if self.ARD else 1)
        lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).copy()
        if self.ARD and lengthscale.size == 1:
            lengthscale = np.ones(self.input_dim) * lengthscale
        if not self.ARD and lengthscale.size != 1:
            raise ValueError("Only 1 lengthscale needed for non-ARD kernel")
        self.lengthscale = lengthscale

    def copy(self):
        return RBF(self.input_dim, self.variance, self.lengthscale.copy(), self.ARD)

    def full_lengthscale(self):
        return self.lengthscale if self.ARD else np.full(self.input_dim, self.lengthscale[0])


class Standardize(object):
    """``GPy.util.normalizer.Standardize`` with the moments reduced on the device / across ranks.
    Nothing is read back while the model is being built: ``mean`` / ``std`` are fetched on first use."""

    def scale_by_device(self, y_dev, cnt_dev):
        """cnt_dev: 1-element device tensor holding this rank's row count; it is summed over ranks
        in the same collective as the first moment (and left holding the global count)."""
        empty = y_dev.numel() == 0                 # a rank without rows still joins both collectives
        s1 = torch.zeros(1, dtype=F64, device=cnt_dev.device) if empty else ops.col_moments(y_dev)[0].clone()
        dist.allreduce_sum_(s1, cnt_dev)
        mean = s1 / cnt_dev
        s2 = torch.zeros(1, dtype=F64, device=cnt_dev.device) if empty else ops.col_moments(y_dev, shift=mean)[1].clone()
        dist.allreduce_sum_(s2)
        self._mean_dev, self._std_dev = mean, torch.sqrt(s2 / cnt_dev)
        self._mean = self._std = None

    def _fetch(self):
        if self._mean is None:
            v = torch.cat([self._mean_dev, self._std_dev]).cpu()
            self._mean, self._std = float(v[0]), float(v[1])

    @property
    def mean(self):
        self._fetch()
        return self._mean

    @mean.setter
    def mean(self, v):
        self._mean = float(v)

    @property
    def std(self):
        self._fetch()
        return self._std

    @std.setter
    def std(self, v):
        self._std = float(v)

    def normalize_device(self, y_dev):
        if y_dev.numel() == 0:
            return y_dev
        return ops.standardize(y_dev, self._mean_dev, self._std_dev)

    def inverse_mean(self, X):
        return (X * self.std) + self.mean

    def inverse_variance(self, var):
        return var * (self.std ** 2)


def _as_device(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=F64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device, non_blocking=True)


_STAGING = {}                 # (device index, doubles) -> [pinned tensor, event of its last copy]


def _upload(a, device):
    """Small host array -> device through a reusable pinned staging buffer, so that the copy is
    asynchronous (a pageable source makes the driver wait for the stream before it returns)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    key = (torch.device(device).index, a.size)
    slot = _STAGING.get(key)
    if slot is None:
        slot = _STAGING[key] = [torch.empty(a.size, dtype=F64, pin_memory=True), None]
    if slot[1] is not None:
        slot[1].synchronize()                      # the previous copy out of this buffer (long done)
    slot[0].numpy()[:] = a.ravel()
    out = slot[0].to(device, non_blocking=True).view(a.shape)
    slot[1] = torch.cuda.Event()
    slot[1].record()
    return out


def _backsub_both_sides(L, X, transpose='left'):
    """GPy.util.linalg.backsub_both_sides on the device: L^-T X L^-1 ('left') or L^-1 X L^-T."""
    trans = transpose == 'left'
    tmp = ops.trsm(L, X.clone(), trans)
    tmp = ops.trsm(L, tmp.t().contiguous(), trans)
    return tmp.t().contiguous()


class NonFiniteInput(ValueError):
    pass


class SparseGPRegression(object):
    """Sparse GP regression (VFE / VarDTC, RBF kernel, Gaussian noise) on one GPU shard.

    Parameters mirror ``GPy.models.SparseGPRegression`` as called from
    ``edrgp/gp_model/regression.py:153-157``.  ``X`` / ``Y`` are this rank's rows (host arrays or
    CUDA tensors); in a multi-process run every rank passes its own shard and the same ``Z`` (or
    ``Z=None`` with the same NumPy seed).
    """

    def __init__(self, X, Y, kernel=None, Z=None, num_inducing=10, X_variance=None, mean_function=None,
                 normalizer=None, device=None, chunk_rows=262144, cache_bytes=None, noise_var=1.0,
                 row_loader=None, pre_sync_check=None, input_dim=None, precision='fp64', y_loader=None):
        if X_variance is not None or mean_function is not None:
            raise NotImplementedError("uncertain inputs / mean functions are outside the B200 path")
        if not torch.cuda.is_available():
            raise RuntimeError("edrgp_b200 needs a CUDA device (there is no CPU fallback)")
        if precision not in ('fp64', 'tf32x3'):
            raise ValueError("precision must be 'fp64' or 'tf32x3'")
        self.precision = precision
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.X = ops.pad_even(_as_device(X, self.device))          # (n_local, d_even)
        if self.X.data_ptr() % 16:                                 # a row slice of a caller's tensor: the kernels
            self.X = self.X.clone()                                # stream rows with 16-byte bulk copies
        self.n_local, self.d_even = self.X.shape
        # 'tf32x3' applies kernel by kernel: cross-covariance and gradients for d <= 64, the weights of the
        # hyper-parameter gradient for m <= 512; everything else runs the FP64 kernels
        # input_dim: X may arrive already padded to an even width (estimator's overlapped loader)
        self.input_dim = X.shape[1] if input_dim is None else int(input_dim)
        Yd = _as_device(Y, self.device).reshape(-1)
        if Yd.shape[0] != self.n_local:
            raise ValueError("X and Y row counts differ")
        if Yd.data_ptr() % 16:
            Yd = Yd.clone()
        # global row count: reduced together with the normaliser's first moment, read back lazily
        self._cnt_dev_t = None
        self._num_data = None if dist.is_distributed() else self.n_local
        if normalizer is True:
            self.normalizer = Standardize()
        elif normalizer is False or normalizer is None:
            self.normalizer = None
        else:
            self.normalizer = normalizer
        # the target moments are reduced lazily: right AFTER the first cross-covariance block has been
        # enqueued (which does not need y), so the device is never idle behind this host-side set-up
        self._Y_raw = Yd
        self._Y_normalized = None
        self._y_loader = y_loader                  # called once, right before y is first touched
        self.kern = RBF(self.input_dim) if kernel is None else kernel
        if self.kern.input_dim != self.input_dim:
            raise ValueError("kernel input_dim does not match X")

        if Z is None:
            self._ensure_y()                         # the draw needs the global row count
            if row_loader is not None:               # the draw gathers arbitrary rows: load them all now
                row_loader(0, self.n_local)
                row_loader = None
            Zh = self._draw_inducing(min(int(num_inducing), self.num_data))
        else:
            Zh = np.array(Z, dtype=np.float64)
            if Zh.ndim != 2 or Zh.shape[1] != self.input_dim:
                raise ValueError("Z must have shape (num_inducing, n_features)")
        self.Z = Zh
        self.num_inducing = Zh.shape[0]
        self.noise_variance = float(noise_var)                     # GPy likelihoods.Gaussian() default: 1.0

        self.chunk_rows = int(max(1024, min(chunk_rows, max(self.n_local, 1))))
        self.chunk_rows += self.chunk_rows & 1                    # even chunks keep y slices 16-byte aligned
        self.cache_bytes = cache_bytes
        self.optimization_runs = []
        self.fix_Z = False
        self._Kcache = None
        self._need_grad = False
        self._fixed = None                    # ops.FixedSweep of the composite fixed-hyper-parameter path
        self._fixed_live = False              # its statistics / alpha belong to the current hyper-parameters
        self._pack_obj = None
        self.kernel_launches = 0
        # row_loader(s, e): called once per row block, right before its first use, to enqueue the
        # host->device copy of rows s:e of X (the estimator overlaps the copy of block i + 1 with the
        # statistics of block i); pre_sync_check(): called before the first host read-back.
        self._row_loader = row_loader
        self._pre_sync_check = pre_sync_check
        self.parameters_changed()
        self._row_loader = None

    _cnt_dev_t = None

    @property
    def _cnt_dev(self):
        """1-element device tensor with this rank's row count (summed over ranks by the normaliser's collective);
        created on first use -- the composite path takes the global count from its own moments table."""
        if self._cnt_dev_t is None:
            self._cnt_dev_t = torch.full((1,), float(self.n_local), dtype=F64, device=self.device)   # (a fill: no H2D)
        return self._cnt_dev_t

    @_cnt_dev.setter
    def _cnt_dev(self, t):
        self._cnt_dev_t = t

    _kuf_flag = None
    _fixed = None
    _fixed_live = False
    _fixed_norm = False
    _pack_obj = None
    _pending = None

    def _ensure_y(self):
        """Normalised targets (GPy ``Standardize``) and the global row count, reduced over ranks."""
        if self._Y_normalized is None:
            if getattr(self, '_y_loader', None) is not None:
                self._y_loader()
                self._y_loader = None
            if self.normalizer is not None:
                self.normalizer.scale_by_device(self._Y_raw, self._cnt_dev)
                self._Y_normalized = self.normalizer.normalize_device(self._Y_raw)
            else:
                dist.allreduce_sum_(self._cnt_dev)
                self._Y_normalized = self._Y_raw
        return self._Y_normalized

    @property
    def Y_normalized(self):
        return self._ensure_y()

    @property
    def num_data(self):
        if self._num_data is None:
            self._ensure_y()
            self._num_data = int(round(float(self._cnt_dev.cpu()[0])))
        return self._num_data

    @num_data.setter
    def num_data(self, v):
        self._num_data = int(v)

    # -------------------------------------------------------------------------------------------
    # inducing inputs: GPy takes ``X[np.random.permutation(n)[:m]]``; with sharded rows every rank
    # draws the same global permutation and contributes the rows it owns.
    # -------------------------------------------------------------------------------------------
    def _draw_inducing(self, m):
        idx = np.random.permutation(self.num_data)[:m]
        lo, _ = dist.shard_bounds(self.num_data) if dist.is_distributed() else (0, self.n_local)
        if dist.is_distributed():
            # shards may be uneven in principle: exchange the true offsets
            sizes = torch.zeros(dist.world_size(), dtype=F64, device=self.device)
            sizes[dist.rank()] = self.n_local
            dist.allreduce_sum_(sizes)
            lo = int(round(float(sizes[:dist.rank()].sum())))
        Zd = torch.zeros(m, self.input_dim, dtype=F64, device=self.device)
        local = idx - lo
        mine = np.nonzero((local >= 0) & (local < self.n_local))[0]
        if mine.size:
            rows = torch.as_tensor(local[mine], device=self.device)
            Zd[torch.as_tensor(mine, device=self.device)] = self.X[rows, :self.input_dim]
        dist.allreduce_sum_(Zd)
        return Zd.cpu().numpy()

    # -------------------------------------------------------------------------------------------
    # pass 1 + solve (+ pass 2)
    # -------------------------------------------------------------------------------------------
    def _chunks(self):
        """Row blocks of the n-scale passes.  While the rows are still arriving from the host
        (``_row_loader``) the blocks are a quarter of ``chunk_rows``: the first kernel then waits for a
        quarter of a block's transfer, and since a block's statistics take longer than the next block's
        copy nothing waits afterwards (growing blocks would stall again at every growth step)."""
        step = self.chunk_rows
        if self._row_loader is not None and self.chunk_rows >= 65536:
            step = (self.chunk_rows // 4) & ~1
        for s in range(0, self.n_local, step):
            yield s, min(self.n_local, s + step)

    def _kbuffers(self, want_cache):
        m = self.num_inducing
        ldk = m + (m & 1)
        if want_cache and self._Kcache is None:
            need = self.n_local * ldk * 8
            budget = self.cache_bytes
            if budget is None:
                # (cudaMemGetInfo costs ~25 ms per call on this driver and the allocator's statistics
                # ~0.2 ms: a matrix under a quarter of the device memory is simply tried)
                total = torch.cuda.get_device_properties(self.device).total_memory
                if need <= total // 4:
                    budget = need
                else:
                    budget = int(0.5 * max(0, total - torch.cuda.memory_allocated(self.device)))
            if need <= budget:
                try:
                    self._Kcache = torch.empty(self.n_local, ldk, dtype=F64, device=self.device)
                except torch.cuda.OutOfMemoryError:
                    self._Kcache = None                       # one reusable block instead (recompute path)
        if self._Kcache is None and (getattr(self, '_Kbuf', None) is None or self._Kbuf.shape[1] != ldk):
            rows = min(self.chunk_rows, self.n_local)                 # one reusable block (recompute path)
            self._Kbuf = torch.empty(rows, ldk, dtype=F64, device=self.device)
        return ldk

    @property
    def _pack(self):
        """Inducing pack with unit coefficients (cross-covariance, Kuu), built on first use."""
        if self._pack_obj is None:
            self._pack_obj = ops.InducingPack(self._Z_dev, self._ell_dev)
        return self._pack_obj

    def _fixed_path_ok(self, need_grad):
        """The composite C-level sweep (``edrgp_fixed_*``) covers the fixed-hyper-parameter evaluation in FP64 for
        even d <= 64 with the rows resident and the whole Kfu kept in HBM."""
        loader = self._row_loader                 # host rows still arriving: fine if the loader exposes the native handle
        return (not need_grad and self.precision == 'fp64' and (loader is None or hasattr(loader, 'handle'))
                and self._Kcache is not None and ops.FixedSweep.supported(self.d_even, self.n_local)
                and (self.normalizer is None or isinstance(self.normalizer, Standardize))
                and self.chunk_rows % 2 == 0)

    def _fixed_pass(self, sf2, beta):
        """Pass 1 + posterior through the composite calls: one C call per stretch between two collectives."""
        m = self.num_inducing
        fs = self._fixed
        # rows still arriving from the host (the estimator's streamer): the calls wait for them block by block, in
        # blocks of the streamer's own size (a quarter of chunk_rows: the first kernel starts after the first block)
        loader = self._row_loader
        chunk = self.chunk_rows if loader is None else min(self.chunk_rows, max(2, loader.rows & ~1))
        h2d, ahead = (loader.handle(), loader.ahead) if loader is not None else (None, 0)
        if fs is None or (fs.n, fs.d, fs.m, fs.chunk, fs.world) != (self.n_local, self.d_even, m, chunk, dist.world_size()):
            fs = self._fixed = ops.FixedSweep.acquire(self.n_local, self.d_even, m, chunk, dist.rank(),
                                                      dist.world_size(), self.device)
            weakref.finalize(self, ops.FixedSweep.release, fs)
            if dist.is_distributed():
                # NVLink peer exchange: the three reductions of the sweep run inside the composite calls (collective
                # set-up on first use per shape; None -> torch.distributed all-reduces between the calls)
                ex = dist.peer_exchange(m, self.d_even)
                if ex is not fs.peer:
                    fs.bind_peers(ex)
        self._y_loader = None                      # (the calls wait for the targets themselves)
        normalize = self._Y_normalized is None and self.normalizer is not None
        y_in = self._Y_raw if self._Y_normalized is None else self._Y_normalized
        fs.begin(self.X, y_in, self._Z_dev, self._ell_dev, sf2, self._Kcache, h2d, ahead)
        if fs.peer is None:
            dist.allreduce_sum_(fs.table)
        fs.stats_pass(self.X, y_in, sf2, self._Kcache, normalize, h2d, ahead)
        if self._Y_normalized is None:
            if normalize:
                norm = self.normalizer
                norm._mean_dev, norm._std_dev = fs.tail[2:3], fs.tail[3:4]
                norm._mean = norm._std = None
                self._Y_normalized = fs.yt
                self._fixed_norm = True
            else:
                self._Y_normalized = self._Y_raw
            self._cnt_dev = fs.tail[1:2]
        if fs.peer is None:
            dist.allreduce_sum_(fs.stats)
        self._stats = (fs.P, fs.byy)                # (peer exchange: summed over the ranks by fs.posterior below)
        fs.posterior(self._Z_dev, sf2, CONST_JITTER, beta)
        self.alpha = fs.alpha
        self._fixed_live = True
        nblk = (self.n_local + chunk - 1) // chunk
        self.kernel_launches += 6 + 3 * nblk + 4 + (m + 31) // 32
        self._enqueue_checks()

    def parameters_changed(self):
        """Recompute the posterior (alpha), the VFE bound and -- while optimising -- its gradient."""
        dev = self.device
        m, d = self.num_inducing, self.input_dim
        sf2 = float(self.kern.variance)
        # lengthscales and inducing inputs travel in ONE pinned upload
        both = np.empty(self.d_even * (m + 1))
        both[:self.d_even] = 1.0
        both[:d] = self.kern.full_lengthscale()
        Zp = both[self.d_even:].reshape(m, self.d_even)
        Zp[:, :d] = self.Z
        if self.d_even != d:
            Zp[:, d:] = 0.0
        ell = both[:self.d_even]
        both_dev = _upload(both, dev)
        self._ell_dev = both_dev[:self.d_even]
        self._Z_dev = both_dev[self.d_even:].view(m, self.d_even)
        beta = 1.0 / max(float(self.noise_variance), CONST_JITTER)
        need_grad = self._need_grad
        # the stored Kfu blocks are reused by the gradient passes when the whole matrix fits
        ldk = self._kbuffers(True)
        self._pack_obj = None
        self._fixed_live = False
        self._beta = beta
        self._solve = None
        self._log_marginal_likelihood = None
        self._woodbury_inv = None
        self._info_dev = None
        self._alpha_is_direct = not need_grad
        if self._fixed_path_ok(need_grad):
            return self._fixed_pass(sf2, beta)

        # TF32-split mode: the training rows' cross-covariance (the n m d contraction) runs on the tcgen05
        # tensor cores; everything downstream (statistics, solve, gradients, eigh) stays FP64
        pack32 = ops.InducingPackTF32(self._Z_dev, self._ell_dev) \
            if (self.precision == 'tf32x3' and self.d_even <= 64) else None
        P = torch.empty(m, m, dtype=F64, device=dev)
        byy = torch.empty(m + 1, dtype=F64, device=dev)
        # the FP64 cross-covariance kernel flags rows holding a NaN / Inf as it forms their norms: the input scan of
        # check_X_y without another pass over X
        self._kuf_flag = torch.zeros(1, dtype=torch.int32, device=dev) if pack32 is None else None
        # statistics route (ops.set_stats_mode): exact INT8 digit products of kernel entries in [0, sf2] (also of the
        # TF32-split mode's entries: they are clipped at sf2 like the FP64 kernel's)
        i8_stats = m <= 2048 and sf2 < 1e150 and ops.get_stats_mode() == 'int8x6'
        for i, (s, e) in enumerate(self._chunks()):
            if self._row_loader is not None:
                self._row_loader(s, e)
            Kc = self._Kcache[s:e] if self._Kcache is not None else self._Kbuf[:e - s]
            if pack32 is not None:
                ops.kuf_tf32(self.X[s:e], pack32, sf2, out=Kc)
            else:
                ops.kuf(self.X[s:e], self._pack, sf2, out=Kc, flag=self._kuf_flag)
            y = self._ensure_y()
            if i8_stats:
                ops.inducing_stats_i8(Kc, y[s:e], sf2, m, P=P, b_yy=byy, accumulate=i > 0)
            else:
                ops.inducing_stats(Kc, y[s:e], m, P=P, b_yy=byy, accumulate=i > 0)
            self.kernel_launches += 4
        if self.n_local == 0:
            self._ensure_y()
            P.zero_(); byy.zero_()
        dist.allreduce_sum_(P, byy)
        self._stats = (P, byy)
        if need_grad:
            trA, data_fit = self._full_chain()
            self._gradients(P, self._solve, beta, sf2, ell[:d], trA, data_fit, float(byy[m]), ldk)
        else:
            # Fixed hyper-parameters: the posterior weights need ONE Cholesky.  With Lm = chol(Kuu) and
            # LB = chol(I + beta Lm^-1 P Lm^-T) of GPy's chain, Lm LB is the Cholesky factor of
            # S = Kuu + beta P, so alpha = Lm^-T LB^-T LB^-1 Lm^-1 beta b = S^-1 beta b.  The bound, the
            # Woodbury inverse and the gradients run the full chain on demand (_ensure_full).
            S = ops.kmm(self._pack, sf2, CONST_JITTER)
            S.add_(P, alpha=beta)
            # one launch per blocked Cholesky step, beta b carried as an extra row (forward solve for free)
            self.alpha, _, info = ops.posv(S, byy[:m] * beta)     # GPy posterior.woodbury_vector (m,)
            self._info_dev = info
            self.kernel_launches += 5 + (m + 31) // 32
            self._enqueue_checks()

    def _full_chain(self):
        """GPy's VarDTC Cholesky chain from the reduced statistics: Lm, LB, alpha, bound."""
        m = self.num_inducing
        P, byy = self._stats
        beta, sf2 = self._beta, float(self.kern.variance)
        Kmm = ops.kmm(self._pack, sf2, CONST_JITTER)
        res = ops.solve(Kmm, P, byy[:m].contiguous(), beta)
        self.kernel_launches += 8
        self._run_pre_sync_check()
        info = res.info.cpu().tolist()
        if info[0] != 0 or info[1] != 0:
            raise np.linalg.LinAlgError("not positive definite: chol(Kuu) info=%d, chol(I + A) info=%d" % tuple(info))
        sc = res.scalars.cpu().numpy()
        trA, sumlogLB, data_fit = float(sc[0]), float(sc[1]), float(sc[2])
        yy = float(byy[m])
        n = self.num_data
        bound = (-0.5 * n * (np.log(2. * np.pi) - np.log(beta)) - 0.5 * beta * yy
                 - 0.5 * (beta * n * sf2 - trA) - sumlogLB + 0.5 * data_fit)
        self._log_marginal_likelihood = np.array([[bound]])
        self._solve = res
        if not self._alpha_is_direct:           # keep one alpha per hyper-parameter set: results stay bit-stable
            self.alpha = res.alpha
        self._info_dev = None
        return trA, data_fit

    def _ensure_full(self):
        if self._solve is None:
            self._full_chain()

    def _enqueue_checks(self):
        """Host-row loader flush + deferred input validation, enqueued (nothing is read back): the
        non-finite count, the Cholesky flag, the row count and the normaliser moments are packed into
        one small device tensor that ``_run_pre_sync_check`` fetches in ONE transfer.  On the composite path
        the kernels have already left all of that in the workspace tail (the cross-covariance kernel flags
        non-finite rows as it forms their norms, the target moments carry a NaN / Inf of y): nothing to enqueue."""
        check, self._pre_sync_check = getattr(self, '_pre_sync_check', None), None
        old = self._pending
        if self._fixed_live:
            on_bad = check(scan=False)[1] if check is not None else None
            if on_bad is None and old is not None:
                on_bad = old[1]
            self._pending = (self._fixed, on_bad, None)
            return
        norm = self.normalizer if isinstance(getattr(self, 'normalizer', None), Standardize) else None
        flag = getattr(self, '_kuf_flag', None)
        if check is None:
            bad, on_bad = None, None
        elif flag is not None and self.n_local > 0:
            # X was scanned by the cross-covariance kernel; the targets show in their own moments (Standardize:
            # checked when the moments are read) or get their own small scan
            on_bad = check(scan=False)[1]
            bad = flag.to(F64) if norm is not None else flag.to(F64) + ops.count_nonfinite(self._Y_raw).to(F64)
        else:
            bad, on_bad = check()
        if getattr(self, '_Y_raw', None) is not None:
            self._ensure_y()
        need_norm = norm is not None and norm._mean is None
        info = getattr(self, '_info_dev', None)
        self._info_dev = None
        if bad is None and info is None and not need_norm and self._num_data is not None:
            return
        if old is not None and isinstance(old[0], torch.Tensor):   # flags of an earlier, still unread evaluation stay armed
            prev, prev_on_bad, _ = old
            bad = prev[0:1] if bad is None else bad.to(F64) + prev[0:1]
            on_bad = on_bad or prev_on_bad
            info = prev[1:2] if info is None else info            # the latest factorisation's flag wins
        if bad is not None and dist.is_distributed():
            bad = bad.to(F64).clone()
            dist.allreduce_sum_(bad)               # every rank raises when any rank's rows are bad
        zero = None
        if bad is None or info is None:
            zero = torch.zeros(1, dtype=F64, device=self.device)
        parts = [bad.to(F64) if bad is not None else zero, info.to(F64) if info is not None else zero, self._cnt_dev]
        if need_norm:
            parts += [norm._mean_dev, norm._std_dev]
        self._pending = (torch.cat(parts), on_bad, norm if need_norm else None)

    def _run_pre_sync_check(self):
        """Fetch what ``_enqueue_checks`` packed (at most ONE read-back, which also brings the row count
        and the normaliser moments to the host); returns the Cholesky flag."""
        self._enqueue_checks()
        pending, self._pending = self._pending, None
        if pending is None:
            return 0
        flat, on_bad, norm = pending
        if isinstance(flat, ops.FixedSweep):
            # composite path: the tail travels behind the eigen-decomposition when that has been read already
            fs = flat
            tail = fs.host[-4:] if fs.host is not None else fs.tail.cpu().numpy()
            flag, info, count, mean, std = ops.FixedSweep.decode_tail(tail)
            if flag & 0x100:
                raise RuntimeError("edrgp_b200: a rank did not reach the peer exchange of the sweep within its time-out "
                                   "(ranks out of step, or a rank died); results of this sweep are invalid")
            self._num_data = int(round(count))
            norm = self.normalizer
            if self._fixed_norm and isinstance(norm, Standardize) and norm._mean is None:
                norm._mean, norm._std = mean, std
            if flag != 0 or (self._fixed_norm and not (np.isfinite(mean) and np.isfinite(std))):
                if on_bad is not None:
                    on_bad()
                raise NonFiniteInput("Input contains NaN or infinity.")
            if info != 0 and fs.world > 1 and not np.isfinite(float(fs.byy[-1])):
                # a NaN / Inf in another rank's rows reaches this rank through the all-reduced statistics
                raise NonFiniteInput("Input contains NaN or infinity (rows of another rank).")
            return info
        v = flat.cpu()
        self._num_data = int(round(float(v[2])))
        bad = float(v[0]) != 0.0
        if norm is not None and norm._mean is None:
            norm._mean, norm._std = float(v[3]), float(v[4])
            bad = bad or not (np.isfinite(norm._mean) and np.isfinite(norm._std))
        if bad and on_bad is not None:
            on_bad()
        return int(v[1])

    def _check_pd(self):
        """Deferred failure check of the sync-free fixed path (called before results reach the host)."""
        info = self._run_pre_sync_check()
        if info != 0:
            raise np.linalg.LinAlgError("not positive definite: chol(Kuu + beta P) info=%d" % info)

    def finish_checks(self):
        """Raise what the deferred checks found (non-finite input: ValueError; Kuu + beta P not
        positive definite: LinAlgError).  One small read-back the first time, nothing afterwards."""
        self._check_pd()
        return self

    def _gradients(self, P, res, beta, sf2, ell, trA, data_fit, yy, ldk):
        dev = self.device
        m, d, n = self.num_inducing, self.input_dim, self.num_data
        # the m x m chain (E, dL_dpsi2, dL_dKmm, sum A o E) in one C call: edrgp_vfe_grad_small
        Mmat, Dsym, sumAE_dev = ops.vfe_grad_small(res, beta)
        self.kernel_launches += 17

        # ---- pass 2 over the rows: T = Kfu o dL/dKfu and its contractions
        S = torch.zeros(m, self.d_even, dtype=F64, device=dev)
        cs = torch.zeros(m, dtype=F64, device=dev)
        mom = torch.zeros(2 * self.d_even, dtype=F64, device=dev)
        if getattr(self, '_Tbuf', None) is None or self._Tbuf.shape[1] != ldk:
            self._Tbuf = torch.empty(min(self.chunk_rows, self.n_local), ldk, dtype=F64, device=dev)
        y = self.Y_normalized
        # TF32-split mode: the n m^2 contraction K M of the weights on tcgen05 (m <= 512)
        tf32_weights = self.precision == 'tf32x3' and m <= 512
        for i, (s, e) in enumerate(self._chunks()):
            if self._Kcache is not None:
                Kc = self._Kcache[s:e]
            else:
                Kc = self._Kbuf[:e - s]
                ops.kuf(self.X[s:e], self._pack, sf2, out=Kc)
                self.kernel_launches += 1
            Tc = self._Tbuf[:e - s]
            if tf32_weights:
                rs = ops.weights_tf32(Kc, Mmat, m, y=y[s:e], alpha=self.alpha, c_ya=beta, c_km=2.0, T=Tc,
                                      want_rowsum=True)
            else:
                rs = ops.weights(Kc, Mmat, m, y=y[s:e], alpha=self.alpha, c_ya=beta, c_km=2.0, T=Tc,
                                 want_rowsum=True, colsum=cs, accumulate=True)
            ops.gemm_tn(Tc, self.X[s:e], ka=m, out=S, accumulate=True)
            ops.col_moments(self.X[s:e], weight=rs, out=mom, accumulate=True)
            self.kernel_launches += 6
        if tf32_weights:
            # column sums of T without a pass over it: sum_i K_ij (beta y_i alpha_j + 2 (K M)_ij)
            #   = beta alpha_j b_j + 2 sum_k P_jk M_kj with the (already reduced) statistics P, b
            dist.allreduce_sum_(S, mom)
            cs = beta * self.alpha * self._stats[1][:m] + 2.0 * (P * Mmat[:, :m]).sum(1)
        else:
            dist.allreduce_sum_(S, cs, mom)
        S = S[:, :d]
        R = mom[self.d_even:self.d_even + d]                     # sum_i rowsum(T)_i x_iq^2
        Z = self._Z_dev[:, :d]
        il2 = torch.as_tensor(1.0 / ell ** 2, device=dev)
        il3 = torch.as_tensor(1.0 / ell ** 3, device=dev)
        sumT = cs.sum()
        gZ = (S - cs[:, None] * Z) * il2
        glen = (R - 2.0 * (Z * S).sum(0) + (cs[:, None] * Z * Z).sum(0)) * il3
        gvar = sumT / sf2

        # ---- Kuu part: T_mm = K(Z, Z) o dL/dKmm (GPy symmetrises through tmp + tmp.T)
        Kzz = ops.kmm(self._pack, sf2, 0.0)
        Tm = Kzz * Dsym
        rs_m = Tm.sum(1)
        TZ = ops.gemm_tn(ops.even_ld(Tm), self._Z_dev, ka=m)[:, :d]         # Tm symmetric: Tm^T Z = Tm Z
        gZ = gZ + 2.0 * (TZ - rs_m[:, None] * Z) * il2
        glen = glen + 2.0 * ((rs_m[:, None] * Z * Z).sum(0) - (Z * TZ).sum(0)) * il3
        gvar = gvar + Tm.sum() / sf2
        self.kernel_launches += 3

        # ---- Kdiag part and the noise (GPy _compute_dL_dR + Gaussian.exact_inference_gradients)
        gvar = float(gvar) - 0.5 * beta * n
        dL_dR = (-0.5 * n * beta + 0.5 * yy * beta ** 2 + 0.5 * (n * sf2 * beta ** 2 - trA * beta)
                 + beta * (0.5 * float(sumAE_dev) - data_fit))
        self.grad_variance = gvar
        glen = glen.cpu().numpy()
        self.grad_lengthscale = glen if self.kern.ARD else np.atleast_1d(glen.sum())
        self.grad_noise = float(dL_dR)
        self.grad_Z = gZ.cpu().numpy()

    def log_likelihood(self):
        """The VFE bound as a (1, 1) array, like GPy (edrgp/tests/test_edr.py:49-50)."""
        self._ensure_full()
        return self._log_marginal_likelihood

    # -------------------------------------------------------------------------------------------
    # paramz-style optimisation: parameters in GPy order (Z, rbf.variance, rbf.lengthscale, noise)
    # -------------------------------------------------------------------------------------------
    def _positive(self):
        return np.concatenate([[self.kern.variance], self.kern.lengthscale, [self.noise_variance]])

    def _get_optimizer_array(self):
        pos = _logexp_finv(self._positive())
        return pos if self.fix_Z else np.concatenate([self.Z.ravel(), pos])

    def _set_optimizer_array(self, x):
        x = np.asarray(x, dtype=np.float64)
        nz = 0 if self.fix_Z else self.Z.size
        if nz:
            self.Z = x[:nz].reshape(self.Z.shape).copy()
        pos = _logexp_f(x[nz:])
        self.kern.variance = float(pos[0])
        self.kern.lengthscale = pos[1:1 + self.kern.lengthscale.size].copy()
        self.noise_variance = float(pos[-1])
        self.parameters_changed()

    def _transformed_gradients(self):
        gpos = np.concatenate([[self.grad_variance], self.grad_lengthscale, [self.grad_noise]])
        gpos = _logexp_gradfactor(self._positive(), gpos)
        return gpos if self.fix_Z else np.concatenate([self.grad_Z.ravel(), gpos])

    _fail_count = 0
    _allowed_failures = 10

    def _objective_grads(self, x):
        fail = 0
        obj, grads = np.inf, None
        try:
            self._set_optimizer_array(x)
            obj = -float(np.sum(self.log_likelihood()))
            grads = -self._transformed_gradients()
        except NonFiniteInput:
            raise                                  # bad input is not a failed evaluation: it surfaces as it is
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError, FloatingPointError):
            fail = 1
        if dist.is_distributed():
            # every rank must walk the same L-BFGS trajectory and stay in the same collectives: a failure on
            # any rank is a failure everywhere, and (f, g) are rank 0's
            buf = torch.zeros(2 + x.size, dtype=F64, device=self.device)
            if not fail:
                buf[1:] = torch.as_tensor(np.concatenate([[obj], grads]), device=self.device)
            buf[0] = float(fail)
            flag = buf[0:1].clone()
            dist.allreduce_max_(flag)
            dist.broadcast_(buf, 0)
            host = buf.cpu().numpy()
            fail = int(round(float(flag.cpu()[0])))
            obj, grads = float(host[1]), host[2:]
        if fail:
            if self._fail_count >= self._allowed_failures:
                raise np.linalg.LinAlgError("the objective could not be evaluated %d times in a row "
                                            "(not positive definite)" % (self._fail_count + 1))
            self._fail_count += 1
            return np.inf, np.clip(np.zeros_like(x), -1e10, 1e10)
        self._fail_count = 0
        return obj, np.clip(grads, -1e10, 1e10)

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, **kwargs):
        """L-BFGS-B on the negative VFE bound (paramz ``Model.optimize`` with its default optimiser)."""
        if optimizer not in (None, 'lbfgsb', 'lbfgs', 'bfgs'):
            raise NotImplementedError("only the default L-BFGS-B optimiser is provided")
        x0 = self._get_optimizer_array() if start is None else np.asarray(start, dtype=np.float64)
        if max_iters <= 0 or x0.size == 0:
            return self
        # deferred input validation surfaces HERE, once and on every rank alike (the non-finite count is summed
        # over ranks), not as a "failed evaluation" inside the optimiser
        self._run_pre_sync_check()
        self._need_grad = True
        try:
            x_opt, f_opt, info = sopt.fmin_l_bfgs_b(self._objective_grads, x0, maxfun=max_iters, maxiter=max_iters)
        finally:
            self._need_grad = False
        self._set_optimizer_array(x_opt)
        self.optimization_runs.append((f_opt, x_opt, info))
        return self

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=False, **kwargs):
        """paramz ``Model.optimize_restarts``: first run from the current point, then random starts."""
        initial = self._get_optimizer_array().copy()
        first = len(self.optimization_runs)
        for i in range(num_restarts):
            try:
                if i > 0:
                    self._set_optimizer_array(np.random.normal(size=initial.size))
                self.optimize(**kwargs)
                if verbose:
                    print("Optimization restart {0}/{1}, f = {2}".format(i + 1, num_restarts,
                                                                         self.optimization_runs[-1][0]))
            except Exception:
                if not robust:
                    raise
        runs = self.optimization_runs[first:]
        if runs:
            best = int(np.argmin([r[0] for r in runs]))
            self._set_optimizer_array(runs[best][1])
        else:
            self._set_optimizer_array(initial)
        return self

    def fixed(self, **kwargs):
        """Keep the current hyper-parameters (``method='fixed'``): the posterior is already in place."""
        return self

    def set_hyperparameters(self, variance=None, lengthscale=None, noise_variance=None, Z=None):
        """Inject hyper-parameters (parity harness and benchmarks run at fixed values)."""
        if variance is not None:
            self.kern.variance = float(variance)
        if lengthscale is not None:
            ls = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).copy()
            if ls.size != self.kern.lengthscale.size:
                raise ValueError("lengthscale has the wrong size")
            self.kern.lengthscale = ls
        if noise_variance is not None:
            self.noise_variance = float(noise_variance)
        if Z is not None:
            Z = np.array(Z, dtype=np.float64)
            if Z.shape != self.Z.shape:
                raise ValueError("Z has the wrong shape")
            self.Z = Z
        self.parameters_changed()
        return self

    # -------------------------------------------------------------------------------------------
    # prediction surface
    # -------------------------------------------------------------------------------------------
    def _grad_scale(self, scale_by_normalizer=True):
        """std(y) of the normaliser (newer GPy scales the Jacobian by it): a float once it has reached
        the host, otherwise the 1-element device tensor, so that nothing is read back mid-sweep."""
        if self.normalizer is not None and scale_by_normalizer:
            norm = self.normalizer
            if isinstance(norm, Standardize) and norm._std is None:
                return norm._std_dev
            return float(norm.std)
        return 1.0

    def _grad_coef(self, scale, factor=1.0):
        """(coef, coef_scale) for an inducing pack carrying alpha * factor * scale."""
        if isinstance(scale, torch.Tensor):
            return self.alpha * scale, float(factor)
        return self.alpha, float(factor) * scale

    def _grad_pack(self, scale):
        coef, cs = self._grad_coef(scale, float(self.kern.variance))
        return ops.InducingPack(self._Z_dev, self._ell_dev, coef, cs)

    def gradient_gram(self, X=None, want_G=False, want_C=True, scale_by_normalizer=True, G_out=None, check=True,
                      reduce=False):
        """Posterior-mean gradients of this rank's rows and their Gram matrix, on the device.

        Returns ``(G or None, C or None)``.  ``reduce=False``: C covers this rank's rows only (callers all-reduce it
        together with whatever else they need).  ``reduce=True``: C is summed over the ranks -- inside the composite
        sweep over NVLink peer memory when the job has exchange buffers (``dist.peer_exchange``), by a
        ``torch.distributed`` all-reduce otherwise; every rank must make the call.  ``X=None`` uses the training
        rows.  ``check=False`` leaves the deferred input / positive-definiteness checks pending (no host
        synchronisation here): the caller runs ``finish_checks()`` before it trusts what it read back.
        """
        Xd = self.X if X is None else ops.pad_even(_as_device(X, self.device))
        scale = self._grad_scale(scale_by_normalizer)
        d = self.input_dim
        reduced = not (reduce and want_C and dist.is_distributed())
        if Xd.shape[0] == 0:
            G = torch.empty(0, d, dtype=F64, device=self.device) if want_G else None
            C = torch.zeros(d, d, dtype=F64, device=self.device) if want_C else None
            if not reduced:
                dist.allreduce_sum_(C)
            return G, C
        use_cache = X is None and getattr(self, '_Kcache', None) is not None
        sf2 = float(self.kern.variance)
        if use_cache and self._fixed_live and self.d_even <= 64 and self.precision == 'fp64':
            # composite path: coefficient pack (std(y) applied on the device), gradients from the stored Kfu and
            # their Gram matrix in one call; C is a view of the result block the eigensolver writes next to
            fs = self._fixed
            G = None
            if want_G:
                G = G_out if (G_out is not None and G_out.shape[1] == self.d_even) else \
                    torch.empty(Xd.shape[0], self.d_even, dtype=F64, device=self.device)
            if isinstance(scale, torch.Tensor):
                C = fs.grad(Xd, self._Kcache, self._Z_dev, self._ell_dev, sf2, 1.0, scale, G)
            else:
                C = fs.grad(Xd, self._Kcache, self._Z_dev, self._ell_dev, sf2, float(scale), None, G)
            if not reduced and fs.peer is not None:
                C = fs.reduce_gram()               # summed over the ranks from the peers' buffers, no collective call
                reduced = True
                self.kernel_launches += 2
            # the caller gets its own copy (it may sum it over ranks in place, keep it, call again); the tag lets
            # GramEighTransformer.fit_gram run the eigensolver inside the block and read everything back at once
            C = C.clone()
            C._edrgp_fixed = weakref.ref(fs)
            self.kernel_launches += 3
        elif use_cache and self.d_even <= 64 and self.precision == 'tf32x3':
            # TF32-split mode: W Z and the row sums as one tcgen05 contraction over the stored Kfu; the
            # Gram matrix of the gradients stays on the FP64 reduction
            coef, cs = self._grad_coef(scale)
            G = ops.grad_tf32(Xd, self._Kcache, self._Z_dev, self._ell_dev, coef, cs, sf2, G_out=G_out)
            C = ops.syrk(G) if want_C else None
            self.kernel_launches += 4
        elif use_cache and self.d_even <= 64:
            # training rows with their cross-covariance already in HBM: no Kuf recompute, no exp
            pack = ops.InducingPack(self._Z_dev, self._ell_dev, *self._grad_coef(scale), block=64)
            G, C = ops.grad_gram_cached(Xd, self._Kcache, pack, sf2, want_G=want_G, want_C=want_C, G_out=G_out)
            self.kernel_launches += 3
        elif not use_cache and self.d_even <= 128:
            fused = self.d_even <= 64
            pack = self._grad_pack(scale)
            G, C = ops.grad_gram(Xd, pack, want_G=want_G or (want_C and not fused), want_C=want_C and fused,
                                 G_out=G_out)
            self.kernel_launches += 3
            if want_C and not fused:
                C = ops.syrk(G)
                self.kernel_launches += 2
        else:
            # any width: row blocks of the stored (or freshly written) Kfu feed the cached-gradient
            # kernel one 64-feature block at a time; the Gram matrix accumulates on the DMMA reduction
            gpack = ops.InducingPack(self._Z_dev, self._ell_dev, *self._grad_coef(scale), block=64)
            kpack = None if use_cache else ops.InducingPack(self._Z_dev, self._ell_dev)
            nrow = Xd.shape[0]
            rows = min(self.chunk_rows, nrow)
            ldk = self.num_inducing + (self.num_inducing & 1)
            Kb = None if use_cache else torch.empty(rows, ldk, dtype=F64, device=self.device)
            G = (G_out if (G_out is not None and G_out.shape[1] == self.d_even) else
                 torch.empty(nrow, self.d_even, dtype=F64, device=self.device)) if want_G else None
            Gb = None if want_G else torch.empty(rows, self.d_even, dtype=F64, device=self.device)
            C = torch.zeros(self.d_even, self.d_even, dtype=F64, device=self.device) if want_C else None
            for i, s0 in enumerate(range(0, nrow, rows)):
                e0 = min(nrow, s0 + rows)
                if use_cache:
                    Kc = self._Kcache[s0:e0]
                else:
                    Kc = Kb[:e0 - s0]
                    ops.kuf(Xd[s0:e0], kpack, sf2, out=Kc)
                Gc = G[s0:e0] if want_G else Gb[:e0 - s0]
                ops.grad_gram_cached(Xd[s0:e0], Kc, gpack, sf2, want_G=True, want_C=False, G_out=Gc)
                if want_C:
                    ops.syrk(Gc, out=C, accumulate=i > 0)
                self.kernel_launches += 6
        if self.d_even != d:
            if G is not None:
                G = G[:, :d].contiguous()
            if C is not None:
                C = C[:d, :d].contiguous()
        if not reduced:
            fixed = getattr(C, '_edrgp_fixed', None)
            dist.allreduce_sum_(C)
            if fixed is not None:
                C._edrgp_fixed = fixed
        if check:
            self._check_pd()
        return (G if want_G else None), C



    def predictive_gradients(self, Xnew, scale_by_normalizer=True):
        """``GP.predictive_gradients``: (mean Jacobian (n, d, 1), None).  edr-gp keeps only
        ``[0][:, :, 0]`` (edrgp/gp_model/base.py:222); the variance gradient GPy also returns (built
        from an n x n matrix and discarded by edr-gp) is not computed.  Newer GPy multiplies the
        Jacobian by std(y) when a normaliser is set; ``scale_by_normalizer`` selects that."""
        G, _ = self.gradient_gram(Xnew, want_G=True, want_C=False, scale_by_normalizer=scale_by_normalizer)
        return G.cpu().numpy()[:, :, None], None

    def woodbury_inv(self):
        """GPy posterior.woodbury_inv = Lm^-T (I - (I + A)^-1) Lm^-1, (m, m) on the device."""
        if self._woodbury_inv is None:
            self._ensure_full()
            m = self.num_inducing
            eye = torch.eye(m, dtype=F64, device=self.device)
            Bi = eye - _backsub_both_sides(self._solve.LB, eye, 'left')
            self._woodbury_inv = _backsub_both_sides(self._solve.Lm, Bi, 'left')
        return self._woodbury_inv

    def predict(self, Xnew, want_variance=True):
        """``GP.predict``: (mean (n, 1), variance (n, 1)) with the likelihood noise added and the
        target normalisation undone."""
        self._check_pd()
        Xd = ops.pad_even(_as_device(Xnew, self.device))
        n = Xd.shape[0]
        m = self.num_inducing
        sf2 = float(self.kern.variance)
        pack = ops.InducingPack(self._Z_dev, self._ell_dev, self.alpha, 1.0)
        mu = torch.empty(n, dtype=F64, device=self.device)
        var = torch.empty(n, dtype=F64, device=self.device) if want_variance else None
        W = ops.even_ld(self.woodbury_inv()) if want_variance else None
        rows = max(2, min(self.chunk_rows, n))
        ldk = m + (m & 1)
        Kb = torch.empty(rows, ldk, dtype=F64, device=self.device) if want_variance else None
        for s in range(0, n, rows):
            e = min(n, s + rows)
            K, _, mu_c = ops.kuf(Xd[s:e], pack, sf2, out=None if Kb is None else Kb[:e - s], want_K=want_variance,
                                 want_mu=True)
            mu[s:e] = mu_c
            if want_variance:
                q = ops.weights(Kb[:e - s], W, m, c_km=1.0, want_rowsum=True)
                var[s:e] = torch.clamp(sf2 - q, min=1e-15) + float(self.noise_variance)
        mu = mu.cpu().numpy()[:, None]
        if var is not None:
            var = var.cpu().numpy()[:, None]
        if self.normalizer is not None:
            mu = self.normalizer.inverse_mean(mu)
            if var is not None:
                var = self.normalizer.inverse_variance(var)
        return mu, var

    # -------------------------------------------------------------------------------------------
    # persistence: plain arrays (edrgp/gp_model/base.py:224-257 pickles the GPy model)
    # -------------------------------------------------------------------------------------------
    def state_dict(self):
        self._ensure_full()
        st = {'Z': self.Z.copy(), 'variance': self.kern.variance, 'lengthscale': self.kern.lengthscale.copy(),
              'ARD': self.kern.ARD, 'noise_variance': self.noise_variance, 'input_dim': self.input_dim,
              'alpha': self.alpha.cpu().numpy(), 'num_data': self.num_data,
              'log_likelihood': float(self._log_marginal_likelihood[0, 0]),
              'Lm': self._solve.Lm.cpu().numpy(), 'LB': self._solve.LB.cpu().numpy()}
        if self.normalizer is not None:
            st['y_mean'], st['y_std'] = self.normalizer.mean, self.normalizer.std
        return st


class FittedSparseGP(SparseGPRegression):
    """A model restored from ``state_dict`` (no training rows): prediction surface only."""

    def __init__(self, state, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("edrgp_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.input_dim = int(state['input_dim'])
        self.d_even = self.input_dim + (self.input_dim & 1)
        self.kern = RBF(self.input_dim, state['variance'], state['lengthscale'], state['ARD'])
        self.Z = np.array(state['Z'], dtype=np.float64)
        self.num_inducing = self.Z.shape[0]
        self.noise_variance = float(state['noise_variance'])
        self.num_data = int(state['num_data'])
        self.n_local = 0
        self.chunk_rows = 262144
        self.kernel_launches = 0
        self.optimization_runs = []
        if 'y_mean' in state:
            self.normalizer = Standardize()
            self.normalizer.mean, self.normalizer.std = float(state['y_mean']), float(state['y_std'])
        else:
            self.normalizer = None
        self._log_marginal_likelihood = np.array([[float(state['log_likelihood'])]])
        ell = np.ones(self.d_even)
        ell[:self.input_dim] = self.kern.full_lengthscale()
        self._ell_dev = torch.as_tensor(ell, device=self.device)
        Zp = np.zeros((self.num_inducing, self.d_even))
        Zp[:, :self.input_dim] = self.Z
        self._Z_dev = torch.as_tensor(Zp, device=self.device)
        self.alpha = torch.as_tensor(np.asarray(state['alpha'], dtype=np.float64), device=self.device)
        sr = ops.SolveResult()
        sr.Lm = torch.as_tensor(np.asarray(state['Lm'], dtype=np.float64), device=self.device)
        sr.LB = torch.as_tensor(np.asarray(state['LB'], dtype=np.float64), device=self.device)
        self._solve = sr
        self._woodbury_inv = None
        self._alpha_is_direct = True
        self.X = None

    def parameters_changed(self):
        raise RuntimeError("a restored model has no training rows; refit to change hyper-parameters")

    def _ensure_full(self):
        pass

    def gradient_gram(self, X=None, **kw):
        if X is None:
            raise ValueError("a restored model has no training rows: pass X")
        return SparseGPRegression.gradient_gram(self, X, **kw)
