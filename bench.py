#!/usr/bin/env python
"""Benchmark of the B200 sparse-GP posterior-gradient EDR path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): EDR fit points/sec at n=4M, d=64, m=512.  One *step* is one
fixed-hyper-parameter EDR sweep over the whole synthetic data set ("full
sweep"): Kfu blocks -> inducing statistics P, b, yy -> [all-reduce] -> m x m Cholesky chain -> alpha
-> posterior-mean gradients -> G^T G -> [all-reduce] -> eigh -> EDR directions.  The n = 4M rows
are sharded over the N ranks (strong scaling: the named shape is the total).

  value  rows / s with the rows already resident in HBM (device tensors passed to the same API)
  e2e    rows / s through the public estimator API with HOST (pinned) buffers: every step copies X
         and y to the device and reads the EDR directions back
  roofline          the dominant kernel (symmetric DMMA reduction P = Kfu^T Kfu) against the FP64
                    tensor-pipe peak measured live by edrgp_fp64_probe
  roofline_pipeline the fused Kuf + gradient + G^T G kernel (the north-star 60 % target)
  cpu_baseline      the NumPy oracle (oracle/pipeline.py, kind "port": GPy is not installable
                    here) timed on the box's host cores on a bounded row sample of the same workload

--impl reference times that CPU port alone, on all host threads, each step a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "EDR fit points/sec at n=4M,d=64,m=512"
UNIT = "points/s"
N_TOTAL, D, M, K_TRUE = 4_000_000, 64, 512, 3
NOISE, SF2 = 0.1, 1.0
# BASELINE.json configs: name -> (n, d, m, chained PCA preprocessor).  The driver's default
# run is C3, the configuration the metric is quoted on; the others are run on request (--config) and their lines are
# kept under profiles/.
CONFIGS = {'C1': (500, 10, 20, False), 'C2': (100_000, 32, 256, False), 'C3': (N_TOTAL, D, M, False),
           'C4': (1_000_000, 512, 1024, False), 'C5': (16_000_000, 128, 2048, True)}
CPU_SAMPLE_ROWS = 131072          # rows per CPU sweep of both CPU legs (cpu_baseline and --impl reference)


def _hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the
    profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs'])
    except (OSError, ValueError, KeyError):
        return 6500.0


HBM_PEAK_GBS = _hbm_peak_gbs()


def flops_per_point(d, m):
    """Algorithmic FP64 flops per point of one sweep (symmetric halves NOT
    discounted): stats 2md + 2m^2 + 3m, pipeline 4md + 2d^2 + m."""
    ntb = (m + 127) // 128
    # executed by the symmetric reduction: full 128 x 128 tiles above the diagonal, and on the
    # diagonal only the 136 upper 8 x 8 blocks of each tile (17/32 of a full tile)
    executed = (ntb * (ntb - 1) // 2) * 2.0 * 128 * 128 + ntb * 136 * 2.0 * 64 + 2.0 * m
    return {'stats': 2.0 * m * d + 2.0 * m * m + 3.0 * m, 'pipeline': 4.0 * m * d + 2.0 * d * d + m,
            'syrk': 2.0 * m * m + 2.0 * m, 'syrk_executed': executed}


def hyper(d, seed=0):
    return np.sqrt(d) * (1. + 0.5 * np.random.RandomState(seed + 1).uniform(size=d))


# ------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def count_since(self, t0):
        return sum(1 for t, _ in list(self.lines) if t >= t0)

    def stop(self, t0=None, t1=None):
        """Clocks over the samples taken from t0 on (perf_counter; None = all); `in_timed_region` counts those up to t1."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = 0
        for ts, ln in self.lines:
            if t0 is not None and ts < t0:
                continue
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            inside += int(t1 is None or ts <= t1)
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "samples_in_timed_region": inside, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_sweep(X, y, Z, ell, sf2, noise, k):
    from oracle import pipeline as op
    n = X.shape[0]
    P, b, yy = op.inducing_stats_chunked(X, y, Z, ell, sf2)
    Kmm = op.kuu(Z, ell, sf2)
    sol = op.solve_from_stats(Kmm, P, b, yy, n, sf2, noise)
    C = op.grad_gram_chunked(X, Z, ell, sf2, sol['alpha'])
    comps, lam, ratio = op.edr_from_gram(C, k)
    return comps


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        info = threadpool_info()
        return max([i.get('num_threads', 1) for i in info] + [1])
    except Exception:
        return os.cpu_count() or 1


def make_cpu_sample(rows, d, m, seed=0):
    from oracle import pipeline as op
    w = op.make_workload(rows, d, m, seed=seed, k_true=K_TRUE)
    w['ell'] = hyper(d, seed)
    return w


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank: lift the BLAS pools back to all host cores (the CPU
    arm runs on rank 0 alone)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def time_cpu(rows, d, m, steps, warmup):
    use_all_host_threads()
    w = make_cpu_sample(rows, d, m)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_sweep(w['X'], w['y'], w['Z'], w['ell'], SF2, NOISE, K_TRUE)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return rows / (sum(times) / len(times)), sum(times) / len(times)


CPU_NOTE = ("CPU port of the reference path (GPy is not installable here): oracle/pipeline.py, row-chunked NumPy / "
            "OpenBLAS in GEMM form on all host threads.  It is FASTER than the reference itself would be: GPy "
            "evaluates the gradient with a per-dimension loop over n x m temporaries and also builds an n x n "
            "variance-gradient term that edr-gp discards, and the reference's SVDTransformer forms an n x n U "
            "(BASELINE.md section 2) -- neither runs at this n.")


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rows = min(args.cpu_rows or CPU_SAMPLE_ROWS, args.n)
    use_all_host_threads()
    cores = cpu_threads()
    pts, sec = time_cpu(rows, args.d, args.m, args.steps, args.warmup)
    sample = ("%d rows of the same generator per step (a bounded sample of the n=%d named shape; throughput in rows/s "
              "is what is compared), all host BLAS threads" % (rows, args.n))
    line = {"impl": "reference", "metric": metric_name(args), "value": pts, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "n": args.n, "d": args.d, "m": args.m,
                       "rows_per_step": rows, "hyperparameters": "fixed", "note": CPU_NOTE},
            "cpu_baseline": {"value": pts, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pts, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def metric_name(args):
    """BASELINE.json's metric string for the configuration it is quoted on (C3); the same metric at the shape
    actually run otherwise."""
    if (args.n, args.d, args.m) == (N_TOTAL, D, M) and not args.pca:
        return METRIC
    return "EDR fit points/sec at n=%d,d=%d,m=%d%s" % (args.n, args.d, args.m, " (PCA-chained)" if args.pca else "")


def workload_name(args):
    name = args.config or next((k for k, v in CONFIGS.items() if v[:3] == (args.n, args.d, args.m)), 'custom')
    return ("%s%s: n=%d d=%d m=%d ARD-RBF sparse-GP EDR sweep at fixed hyper-parameters (%sstats + solve + posterior "
            "gradients + GtG + eigh), rows sharded over ranks"
            % (name, " headline" if name == 'C3' else "", args.n, args.d, args.m,
               "StandardScaler + PCA preprocessor + " if args.pca else ""))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as tdist
    import edrgp_b200 as eb
    from edrgp_b200 import dist as edist, model as emodel, ops

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # multi-rank runs: keep each rank (and the pinned host buffers it allocates from here on) on the CPUs next
    # to its GPU; a single rank keeps every host core (the CPU baseline runs in this process)
    host_affinity = edist.bind_to_local_cpus(local_rank) if world > 1 else 'unchanged (single rank)'
    # ONE JSON line on stdout: native libraries that write to file descriptor 1 (NCCL's version banner)
    # are sent to stderr; the line itself goes through a duplicate of the original descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        tdist.init_process_group('nccl', device_id=dev)
    n, d, m = args.n, args.d, args.m
    ops.set_stats_mode(args.stats)
    lo, hi = edist.shard_bounds(n, rank, world)
    n_local = hi - lo

    # ---- synthetic data of the named shape: X ~ N(0, I) standardised,
    # y = sum tanh(X B) + 0.05 eps, B (d x 3) orthonormal; Z = m random rows; fixed ARD lengthscales.
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.randn(n_local, d, dtype=torch.float64, device=dev, generator=g)
    Bm = torch.as_tensor(np.linalg.qr(np.random.RandomState(0).standard_normal((d, K_TRUE)))[0], device=dev)
    y = torch.tanh(X @ Bm).sum(1) + 0.05 * torch.randn(n_local, dtype=torch.float64, device=dev, generator=g)
    Z0 = X[:m].clone() if rank == 0 else torch.zeros(m, d, dtype=torch.float64, device=dev)
    if world > 1:
        tdist.broadcast(Z0, 0)
    Z = Z0.cpu().numpy()
    ell = hyper(d)

    def make_estimator(precision=None):
        return eb.SparseGaussianProcessRegressor(kernels=emodel.RBF(d, SF2, ell, ARD=True), Z=Z, normalizer=True,
                                                 method='fixed', noise_var=NOISE, chunk_rows=args.chunk_rows,
                                                 deferred_checks=True, precision=precision or args.precision)

    def preprocess(Xin):
        """C5: the front of the chain on the device (edrgp/edr.py:142-176 with preprocessor=PCA): column moments
        + standardise, centred Gram matrix on the DMMA reduction + eigh + projection (DevicePCA), all-reduced."""
        Xd = Xin if isinstance(Xin, torch.Tensor) else torch.as_tensor(Xin, device=dev)
        cnt = torch.tensor([float(Xd.shape[0])], dtype=torch.float64, device=dev)
        s1 = ops.col_moments(Xd)[0].clone()
        edist.allreduce_sum_(s1, cnt)
        mean = s1 / cnt
        s2 = ops.col_moments(Xd, shift=mean)[1].clone()
        edist.allreduce_sum_(s2)
        Xs = ops.standardize(Xd, mean, torch.sqrt(s2 / cnt))
        return eb.DevicePCA(n_components=d).fit_transform_device(Xs)

    def sweep(Xin, yin, precision=None):
        """One fixed-hyper-parameter EDR sweep through the public classes.  The input validation and
        the Cholesky flag are computed on the device inside the sweep and read (and raised) once, after
        the directions have been read back: no host round trip between the row passes."""
        if args.pca:
            Xin = preprocess(Xin)
            if not isinstance(yin, torch.Tensor):
                yin = torch.as_tensor(yin, device=dev)
        est = make_estimator(precision).fit(Xin, yin)
        _, C = est.estimator_.gradient_gram(want_G=False, check=False, reduce=True)     # summed over the ranks
        tr = eb.GramEighTransformer(n_components=K_TRUE).fit_gram(C, n)
        est.estimator_.finish_checks()
        return tr.components_

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms[0]), out

    # ---- device-resident arm.  The clock sampler (nvidia-smi -lms 100) is started BEFORE the warm-up, so that it is up
    # and printing when the timed region begins: on 8 GPUs that region is under 0.1 s.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        comps = sweep(X, y)

    # ---- FP64 tensor-pipe peak (denominator of the roofline): live probe on the warm device,
    # cross-checked against the standalone microbenchmark recorded under profiles/
    peak_live = ops.fp64_tensor_peak_tflops()
    peak_recorded = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'fp64_peak_r01.json')) as f:
            peak_recorded = float(json.load(f)['fp64_dmma_tflops'])
    except (OSError, ValueError, KeyError):
        pass
    peak = max(peak_live, peak_recorded or 0.0)
    launches0 = ops.launch_count()
    ops.start_timing()
    t_region0 = time.perf_counter()
    total_ms, comps = timed(lambda: sweep(X, y), args.steps)
    t_region1 = time.perf_counter()
    per_op = ops.stop_timing()
    launches = ops.launch_count() - launches0
    # a timed region shorter than a few sampling periods: the same sweeps keep the device under the same load (untimed)
    # until the sampler has three readings since the region began -- every rank runs them, rank 0 decides
    extra_sweeps = 0
    while True:
        need = torch.tensor([1.0 if (rank == 0 and sampler.proc is not None and sampler.count_since(t_region0) < 3
                                     and extra_sweeps < 400) else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.broadcast(need, 0)
        if float(need[0]) == 0.0:
            break
        sweep(X, y)
        extra_sweeps += 1
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    if clocks is not None and extra_sweeps:
        clocks["untimed_sweeps_under_the_sampler_after_the_region"] = extra_sweeps
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # ---- the fused Kuf + gradient + GtG kernel on its own (north-star pipeline roofline).  The sweep
    # above feeds the gradient pass from the Kfu blocks of the statistics pass when they fit in HBM;
    # the recompute kernel is what runs for new rows and when they do not.
    pipe_alone_ms, pipe_alone_n = 0.0, 1
    if d + (d & 1) <= 64 and not args.pca:
        est0 = make_estimator().fit(X, y)
        gpack = est0.estimator_._grad_pack(1.0)
        for _ in range(2):
            ops.grad_gram(est0.estimator_.X, gpack, want_G=False)
        ops.start_timing()
        for _ in range(3):
            ops.grad_gram(est0.estimator_.X, gpack, want_G=False)
        pa = ops.stop_timing()
        pipe_alone_ms, pipe_alone_n = pa.get('grad_gram', (0.0, 1))
        del est0, gpack

    # ---- the same sweep in the TF32-split mode (informational; the headline above is FP64): the training
    # rows' cross-covariance on the tcgen05 tensor cores, everything downstream unchanged
    tf32_mode = None
    if d + (d & 1) <= 64 and args.precision == 'fp64' and not args.no_tf32 and not args.pca:
        for _ in range(2):
            sweep(X, y, 'tf32x3')
        ops.start_timing()
        t_steps = max(1, min(args.steps, 5))
        t_ms, comps32 = timed(lambda: sweep(X, y, 'tf32x3'), t_steps)
        t_ops = ops.stop_timing()
        # ... and with the statistics on the INT8 tensor cores as well (both tcgen05 routes together)
        both_ms = None
        if m <= 2048 and not args.no_int8:
            ops.set_stats_mode('int8x6')
            try:
                for _ in range(2):
                    sweep(X, y, 'tf32x3')
                both_ms, comps_both = timed(lambda: sweep(X, y, 'tf32x3'), t_steps)
            finally:
                ops.set_stats_mode(args.stats)
        k_ms, k_n = t_ops.get('kuf', (0.0, 1))
        rows_chk = min(n_local, 4096)
        pk64 = ops.InducingPack(torch.as_tensor(Z, device=dev), torch.as_tensor(ell, device=dev))
        pk32 = ops.InducingPackTF32(torch.as_tensor(Z, device=dev), torch.as_tensor(ell, device=dev))
        K64, _ = ops.kuf(X[:rows_chk], pk64, SF2)
        K32 = ops.kuf_tf32(X[:rows_chk], pk32, SF2)
        from edrgp_b200.utils import principal_angle as _pa
        tf32_mode = {
            "ms_per_step": t_ms / t_steps, "value": n / (t_ms / t_steps * 1e-3), "unit": UNIT, "steps": t_steps,
            "kernel": "kuf_tf32_kernel (tcgen05.mma kind::tf32, 3 products per entry, FP32 accumulators in TMEM)",
            "kuf_ms_per_step": k_ms / t_steps, "kuf_avg_launch_ms": k_ms / max(k_n, 1),
            "with_int8x6_stats": None if both_ms is None else {
                "ms_per_step": both_ms / t_steps, "value": n / (both_ms / t_steps * 1e-3), "unit": UNIT,
                "leading_direction_angle_vs_fp64_rad": float(__import__('edrgp_b200').utils.principal_angle(comps_both[:1], comps[:1]))},
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": HBM_PEAK_GBS,
                         "achieved": (m + d) * 8.0 * n_local * t_steps / (k_ms * 1e-3) / 1e9 if k_ms else None,
                         "frac": (m + d) * 8.0 * n_local * t_steps / (k_ms * 1e-3) / 1e9 / HBM_PEAK_GBS if k_ms else None,
                         "note": "algorithmic bytes per point: 8 m written (Kfu, FP64) + 8 d read (X); peak = "
                                 "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"},
            "max_rel_err_entries_vs_fp64": float(((K32 - K64).abs() / K64).max()),
            "leading_direction_angle_vs_fp64_rad": float(_pa(comps32[:1], comps[:1])),
            "tolerance": "1e-4 relative on kernel entries (BASELINE north_star, TF32-split mode)"}
        del K64, K32, pk64, pk32

    # ---- the same sweep with the statistics on the INT8 tensor cores (informational unless --stats int8x6 made it
    # the headline): exact integer products of six 8-bit slices of Kfu, everything else unchanged
    int8_mode = None
    if args.precision == 'fp64' and args.stats == 'fp64' and not args.no_int8 and m <= 2048:
        ops.set_stats_mode('int8x6')
        try:
            for _ in range(3):
                sweep(X, y)
            ops.start_timing()
            i_steps = max(1, min(args.steps, 5))
            i_ms, comps8 = timed(lambda: sweep(X, y), i_steps)
            i_ops = ops.stop_timing()
            s_ms, s_n = i_ops.get('inducing_stats', (0.0, 1))
            from edrgp_b200.utils import principal_angle as _pa8
            int8_mode = {
                "ms_per_step": i_ms / i_steps, "value": n / (i_ms / i_steps * 1e-3), "unit": UNIT, "steps": i_steps,
                "kernels": "slice_u8_kernel + syrk_i8_kernel (tcgen05.mma kind::i8, 21 slice pairs, int32 accumulators "
                           "in TMEM) + i8_reduce_kernel",
                "inducing_stats_ms_per_step": s_ms / i_steps,
                "max_abs_diff_components_vs_fp64": float(np.max(np.abs(np.abs(comps8) - np.abs(comps)))),
                "leading_direction_angle_vs_fp64_rad": float(_pa8(comps8[:1], comps[:1])),
                "how": "ops.set_stats_mode('int8x6') / EDRGP_STATS=int8x6 / bench.py --stats int8x6"}
        finally:
            ops.set_stats_mode('fp64')

    # quality of the directions found (not a timing): principal angle to the true subspace
    from edrgp_b200.utils import principal_angle
    Bt = Bm.cpu().numpy().T
    lead = comps[0] / np.linalg.norm(comps[0])
    angle_lead = float(np.arcsin(min(1.0, np.linalg.norm(lead - Bt.T.dot(Bt.dot(lead))))))
    angle = principal_angle(comps, Bt)

    # ---- end-to-end arm: host (pinned) rows in, directions out, every step
    Xh = torch.empty(n_local, d, dtype=torch.float64, pin_memory=True)
    yh = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    Xh.copy_(X); yh.copy_(y)
    torch.cuda.synchronize()
    Xnp, ynp = Xh.numpy(), yh.numpy()
    del X, y
    torch.cuda.empty_cache()

    def sweep_host():
        return sweep(Xnp, ynp)          # host arrays straight into the public API

    for _ in range(2):
        sweep_host()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_ms, comps_h = timed(sweep_host, e2e_steps)
    e2e_value = n / (e2e_ms / e2e_steps * 1e-3)
    h2d = (n * d + n) * 8
    d2h = (K_TRUE * d + d + d * d) * 8 + 64

    # ---- the same end-to-end arm from PAGEABLE NumPy arrays (what an sklearn user passes): the estimator stages
    # them through its ring of pinned blocks (native host threads), one cudaMemcpyAsync per block
    e2e_pageable = None
    if not args.no_pageable:
        Xpg, ypg = np.array(Xnp), np.array(ynp)             # plain malloc'ed copies
        for _ in range(2):
            sweep(Xpg, ypg)
        pg_steps = max(1, min(args.steps, 5))
        pg_ms, _ = timed(lambda: sweep(Xpg, ypg), pg_steps)
        e2e_pageable = {"value": n / (pg_ms / pg_steps * 1e-3), "unit": UNIT, "ms_per_step": pg_ms / pg_steps,
                        "steps": pg_steps, "ratio_to_pinned": (e2e_ms / e2e_steps) / (pg_ms / pg_steps),
                        "note": "host rows in ordinary (pageable) NumPy arrays"}
        del Xpg, ypg

    # ---- the public orchestrator with the reference's full semantics (informational, N = 1 only):
    # EffectiveDimensionalityReduction.fit = StandardScaler + GP fit + gradients + eigh + projection +
    # the second GP fit on the projected rows (edrgp/base.py:172-200), from the same host rows
    edr_fit = None
    if world == 1 and not args.no_edr:
        def edr_full():
            return eb.EffectiveDimensionalityReduction(make_estimator(), eb.GramEighTransformer(), n_components=None,
                                                       normalize=True, keep_gradients=False).fit(Xnp, ynp)
        for _ in range(2):
            edr_full()
        edr_ms, edr_obj = timed(edr_full, 3)
        edr_fit = {"ms_per_fit": edr_ms / 3, "value": n / (edr_ms / 3 * 1e-3), "unit": UNIT,
                   "what": "EffectiveDimensionalityReduction(estimator, GramEighTransformer(), n_components=None)"
                           ".fit(X_host, y_host): scaler + sweep + projection + last fit (two GP fits), H2D included",
                   "num_iter": int(edr_obj.num_iter)}

    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return

    fl = flops_per_point(d, m)
    syrk_ms, syrk_n = per_op.get('inducing_stats', (0.0, 1))
    cached_ms, cached_n = per_op.get('grad_gram_cached', (0.0, 0))
    pipe_ms, pipe_n = pipe_alone_ms, pipe_alone_n
    kuf_ms, _ = per_op.get('kuf', (0.0, 1))
    solve_ms, _ = per_op.get('solve', (0.0, 1))
    eigh_ms, _ = per_op.get('eigh', (0.0, 1))
    steps = args.steps
    rows_per_launch_syrk = n_local * steps / max(syrk_n, 1)
    syrk_avg_ms = syrk_ms / max(syrk_n, 1)
    syrk_tf = fl['syrk'] * rows_per_launch_syrk / (syrk_avg_ms * 1e-3) / 1e12 if syrk_avg_ms else 0.0
    syrk_tf_exec = fl['syrk_executed'] * rows_per_launch_syrk / (syrk_avg_ms * 1e-3) / 1e12 if syrk_avg_ms else 0.0
    pipe_avg_ms = pipe_ms / max(pipe_n, 1)
    pipe_tf = fl['pipeline'] * n_local / (pipe_avg_ms * 1e-3) / 1e12 if pipe_avg_ms else 0.0

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        rows = min(args.cpu_rows or CPU_SAMPLE_ROWS, n)
        cpu_steps = 5
        pts, sec = time_cpu(rows, d, m, cpu_steps, 1)
        cpu = {"value": pts, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": "%d rows of the same generator per sweep (the sample --impl reference uses), %d sweeps after "
                         "one warm-up (%.2f s each); oracle/pipeline.py chunked NumPy/OpenBLAS on all host threads"
                         % (rows, cpu_steps, sec), "note": CPU_NOTE}

    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of a launch of the
    # same size (524288 rows); null when the launch size differs or the capture is absent
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')) as f:
            tr = json.load(f)['gemm_tn_kernel']
        if abs(rows_per_launch_syrk - tr['rows']) < 0.1 * tr['rows'] and (d, m) == (D, M):
            traffic = tr['dram_bytes_read'] + tr['dram_bytes_write']
    except (OSError, ValueError, KeyError):
        pass

    if world == 1:
        collectives = "single rank: none"
    elif d <= 64 and edist.peer_exchange(m, d + (d & 1)) is not None:
        collectives = ("NVLink peer exchange inside the library's kernels: moments table pushed, {P, b, yy} summed while the "
                       "system is assembled, C summed in front of eigh; no library collective on the sweep")
    else:
        collectives = "torch.distributed all-reduce over NCCL"
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.precision == 'fp64' else "f64 (cross-covariance contraction: tf32 x 3)", "data": "synthetic",
        "config": {"workload": workload_name(args), "n": n, "d": d, "m": m, "rows_per_rank": n_local,
                   "host_affinity": host_affinity,
                   "hyperparameters": "fixed (lengthscales sqrt(d)(1+u/2), variance 1, noise 0.1)",
                   "l2": "inputs (%.2f GB X per rank + %.1f GB Kfu blocks) exceed the 126 MB L2; no flush needed"
                         % (n_local * d * 8 / 1e9, n_local * m * 8 / 1e9),
                   "chunk_rows": args.chunk_rows, "precision": args.precision, "stats": args.stats, "parallelism": "n-sharded x%d, 3 small reductions/step (%s)" % (world, collectives)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "api": "SparseGaussianProcessRegressor(method='fixed').fit(X_host, y_host) -> gradient_gram -> "
                       "GramEighTransformer.fit_gram -> components_ (host)",
                "host_memory": "pinned", "pageable": e2e_pageable},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_tn_kernel (P = Kfu^T Kfu, b, yy; symmetric DMMA reduction)",
                     "achieved": syrk_tf_exec, "peak": peak, "unit": "TFLOP/s", "frac": syrk_tf_exec / peak if peak else None,
                     "flops_counted": "EXECUTED by the kernel: the upper triangle of P only (full 128 x 128 tiles above "
                                      "the diagonal, 8 x 8 blocks on it) = %.1f %% of the algorithmic 2 m^2 + 2 m per row; "
                                      "this is the pipe-utilisation figure ncu's sm__pipe_tensor_subpipe_dmma reports"
                                      % (100.0 * fl['syrk_executed'] / fl['syrk']),
                     "achieved_algorithmic": syrk_tf, "frac_algorithmic": syrk_tf / peak if peak else None,
                     "algorithmic_note": "SURVEY 8d counts 2 m^2 + 2 m flops per row with the symmetric half NOT "
                                         "discounted; against that count the kernel reads above 1 by construction",
                     "traffic": traffic,
                     "traffic_source": "recorded: dram__bytes_read.sum + dram__bytes_write.sum of one 524288-row launch "
                                       "(ncu --set full, profiles/r01_ncu_top_kernels_final.txt); null when this run's "
                                       "launch size differs.  Algorithmic bytes per launch = 2.15 GB (the Kfu block "
                                       "once): each column slab is read by several tiles, L2 absorbs about half",
                     "peak_live": peak_live, "peak_recorded": peak_recorded,
                     "peak_source": "measured FP64 DMMA (mma.sync m8n8k4 f64) peak of this pool's B200: the larger of "
                                    "the live probe in this process (edrgp_fp64_probe, CUDA events) and the standalone "
                                    "microbenchmark recorded in profiles/fp64_peak_r01.json; MEASURED_PEAKS.json has "
                                    "no FP64 figure",
                     "avg_launch_ms": syrk_avg_ms, "launches": syrk_n, "share_of_step": syrk_ms / total_ms},
        "roofline_pipeline": {"bound": "tensor", "kernel": "grad_gram_kernel (fused Kuf + gradient + GtG)",
                              "achieved": pipe_tf, "peak": peak, "unit": "TFLOP/s",
                              "frac": pipe_tf / peak if peak else None, "avg_launch_ms": pipe_avg_ms,
                              "launches": pipe_n, "share_of_step": None,
                              "note": "timed on its own over the same rows right after the timed region (3 launches); "
                                      "inside the sweep the gradient pass reads the stored Kfu blocks instead "
                                      "(grad_gram_cached) when the whole matrix fits in HBM",
                              "algorithmic_flops_per_point": fl['pipeline'], "exps_per_point": m},
        "stage_ms_per_step": {"kuf": kuf_ms / steps, "inducing_stats": syrk_ms / steps, "solve": solve_ms / steps,
                              "grad_gram_cached": cached_ms / steps,
                              "grad_gram_recompute": per_op.get('grad_gram', (0.0, 0))[0] / steps,
                              "eigh": eigh_ms / steps},
        "roofline_grad_cached": {"bound": "tensor", "kernel": "grad_gram_kernel<FROM_K> (stored Kfu -> W Z, row sums, GtG)",
                                 "achieved": (2.0 * m * d + 2.0 * d * d + 2.0 * m) * n_local * steps
                                 / (cached_ms * 1e-3) / 1e12 if cached_ms else None,
                                 "peak": peak, "unit": "TFLOP/s",
                                 "hbm_gbs": (m + d + 1) * 8.0 * n_local * steps / (cached_ms * 1e-3) / 1e9 if cached_ms else None,
                                 "avg_launch_ms": cached_ms / max(cached_n, 1), "launches": cached_n},
        "sweep_tflops_algorithmic": (fl['stats'] + fl['pipeline']) * n / (ms_per_step * 1e-3) / 1e12,
        "quality": {"leading_direction_angle_to_true_subspace_rad": angle_lead,
                    "largest_principal_angle_k3_rad": angle,
                    "note": "fixed, un-optimised hyper-parameters; y = sum tanh(x.b_k) is dominated by one "
                            "direction, so only the leading direction is expected to lie in span(B)"},
        "cpu_baseline": cpu,
        "edr_fit": edr_fit,
        "tf32x3_mode": tf32_mode,
        "int8x6_stats_mode": int8_mode,
    }
    if args.stats == 'int8x6':
        # the headline ran with the statistics on the INT8 tensor cores: describe THAT stage (digit pass + reduction +
        # split sum are bracketed together by the library's stage timer)
        nt = (m + 127) // 128
        ops_row = 21 * 2.0 * (nt * (nt + 1) // 2) * 128 * 128           # executed: 21 digit pairs on the upper tiles
        tops = ops_row * rows_per_launch_syrk / (syrk_avg_ms * 1e-3) / 1e12 if syrk_avg_ms else 0.0
        line["dtype"] = "f64 (inducing statistics: exact int8 digit products, int32 accumulators)"
        line["roofline"] = {
            "bound": "tensor", "kernel": "slice_u8_kernel + syrk_i8_kernel + i8_reduce_kernel (the statistics stage per row block)",
            "achieved": tops, "peak": 4500.0, "unit": "TOP/s", "frac": tops / 4500.0,
            "peak_source": "NOMINAL dense INT8 rate of a B200 (4.5 POP/s): MEASURED_PEAKS.json has no INT8 figure",
            "ops_counted": "EXECUTED by syrk_i8_kernel: 21 digit pairs x 2 x 128 x 128 per row on the %d upper 128 x 128 "
                           "tiles; the stage time also contains the HBM-bound digit pass (about 40 %% of it)" % (nt * (nt + 1) // 2),
            "achieved_fp64_equivalent_tflops": syrk_tf, "fp64_dmma_peak_tflops": peak,
            "fp64_equivalent_note": "the algorithmic 2 m^2 + 2 m FP64 flops per row this stage replaces, over its time",
            "traffic": None, "traffic_source": "profiles/r02_ncu_top_kernels.txt (per kernel)",
            "avg_launch_ms": syrk_avg_ms, "launches": syrk_n, "share_of_step": syrk_ms / total_ms}
    print(json.dumps(line), file=real_stdout)
    real_stdout.flush()
    if world > 1:
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', type=int, default=None, help="rows (default: the configuration's)")
    ap.add_argument('--d', type=int, default=None)
    ap.add_argument('--m', type=int, default=None)
    ap.add_argument('--chunk-rows', type=int, default=524288)
    ap.add_argument('--cpu-rows', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-edr', action='store_true')
    ap.add_argument('--no-tf32', action='store_true', help="skip the informational TF32-split arm")
    ap.add_argument('--no-int8', action='store_true', help="skip the informational INT8-slice statistics arm")
    ap.add_argument('--stats', default='fp64', choices=['fp64', 'int8x6'],
                    help="statistics route of the sweep (the headline is fp64: FP64 DMMA)")
    ap.add_argument('--no-pageable', action='store_true', help="skip the pageable-host-memory end-to-end leg")
    ap.add_argument('--config', default=None, choices=sorted(CONFIGS), help="a BASELINE.json configuration (default: C3)")
    ap.add_argument('--pca', action='store_true', help="run the StandardScaler + DevicePCA front of the chain inside the step")
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'tf32x3'],
                    help="arithmetic of the cross-covariance pass (the headline is fp64)")
    args = ap.parse_args()
    cn, cd, cm, pca = CONFIGS[args.config or 'C3']
    args.n, args.d, args.m = args.n or cn, args.d or cd, args.m or cm      # explicit sizes override the configuration's
    args.pca = args.pca or (pca and args.config is not None)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
