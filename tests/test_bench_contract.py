"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
(`--impl reference`: the CPU port of the reference path timed on the host cores) prints exactly one JSON
line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-rows', '8192'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'].startswith('EDR fit points/sec') and d['unit'] == 'points/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_default_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must exit non-zero instead of printing a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, 'bench.py', '--steps', '1', '--warmup', '0', '--no-cpu'], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith('{') and '"value"' in l]


def test_config_selection_and_metric_names():
    """--config picks a BASELINE.json shape, explicit sizes override it, and only the headline shape carries
    BASELINE's metric string."""
    import importlib
    import types
    bench = importlib.import_module('bench')
    def args(**kw):
        a = types.SimpleNamespace(config=None, n=None, d=None, m=None, pca=False)
        a.__dict__.update(kw)
        cn, cd, cm, pca = bench.CONFIGS[a.config or 'C3']
        a.n, a.d, a.m = a.n or cn, a.d or cd, a.m or cm
        a.pca = a.pca or (pca and a.config is not None)
        return a
    assert bench.metric_name(args()) == bench.METRIC
    a = args(config='C4')
    assert (a.n, a.d, a.m, a.pca) == (1_000_000, 512, 1024, False) and 'n=1000000,d=512,m=1024' in bench.metric_name(a)
    a = args(config='C5', n=2_000_000)
    assert (a.n, a.d, a.m, a.pca) == (2_000_000, 128, 2048, True) and 'PCA-chained' in bench.metric_name(a)
    assert 'C5' in bench.workload_name(a) and 'PCA preprocessor' in bench.workload_name(a)
    fl = bench.flops_per_point(64, 512)
    assert fl['syrk'] == 2 * 512 * 512 + 2 * 512 and 0.5 < fl['syrk_executed'] / fl['syrk'] < 0.52
