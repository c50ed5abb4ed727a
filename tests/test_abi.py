"""The C-ABI library loads and exports every symbol include/edrgp_b200.h declares, and the ctypes
table mirrors the header one to one.  No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'edrgp_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(edrgp_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
    names = _declared()
    for must in ('edrgp_kuf', 'edrgp_grad_gram', 'edrgp_inducing_stats', 'edrgp_solve', 'edrgp_eigh',
                 'edrgp_weights', 'edrgp_gemm_tn', 'edrgp_project', 'edrgp_col_moments'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from edrgp_b200 import build, _lib
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), name


def test_ctypes_table_mirrors_header():
    from edrgp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r'\b%s\s*\(([^;]*)\)\s*;' % name, src)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else len(params.split(','))
        assert n == len(args), (name, n, len(args))


def test_version_and_error_calls_work_without_gpu():
    from edrgp_b200 import _lib
    lib = _lib.load()
    assert lib.edrgp_version() >= 100
    assert lib.edrgp_pack_bytes(512, 64) == (64 + 16 * 32 * (64 + 2 + 2)) * 8
    # argument errors are reported before any CUDA call
    rc = lib.edrgp_kuf(0, 4, 10, 4, 0, 3, 1.0, 0, 4, 0, 0, 0, 0, 0, 0)
    assert rc == -1
    assert b'kuf' in lib.edrgp_last_error()


def test_product_path_fails_loudly_without_cuda():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import edrgp_b200 as eb
    with pytest.raises(Exception):
        eb.SparseGaussianProcessRegressor(method='fixed').fit(np.zeros((10, 2)), np.zeros(10))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'edrgp_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn


def test_fixed_sweep_layout_and_tail_decoding():
    """The composite sweep's workspace layout (no GPU needed: sizes only) and the packing of the two 32-bit check
    words into the first tail double."""
    import ctypes
    import numpy as np
    from edrgp_b200 import _lib, ops
    lib = _lib.load()
    names = ops.FixedSweep.REGIONS
    off = (ctypes.c_int64 * len(names))()
    nbytes = lib.edrgp_fixed_layout(500000, 64, 512, 524288, 8, off)
    assert nbytes > 0 and nbytes % 16 == 0
    o = dict(zip(names, list(off)))
    assert all(v % 2 == 0 for v in o.values())                       # 16-byte aligned regions
    order = sorted(o.values())
    assert order == [o[k] for k in names]                              # laid out in declaration order
    assert o['stats'] - o['yt'] >= 500000 and o['table'] - o['stats'] >= 512 * 512 + 512 + 1
    assert o['S'] - o['table'] >= 4 * 8
    assert nbytes // 8 - o['result'] >= 64 + 2 * 64 * 64 + 4
    assert lib.edrgp_fixed_layout(0, 64, 512, 524288, 1, off) == 0    # rejected shapes
    tail = np.zeros(4)
    tail[:1] = np.array([3, 41], dtype=np.int32).view(np.float64)
    tail[1:] = [4e6, 0.25, 1.75]
    assert ops.FixedSweep.decode_tail(tail) == (3, 41, 4e6, 0.25, 1.75)
    assert ops.FixedSweep.supported(64, 10) and not ops.FixedSweep.supported(66, 10) and not ops.FixedSweep.supported(64, 0)


def test_stats_mode_switch_and_layout():
    """The statistics route is process-wide state of the library (no GPU needed to flip it): the sweep workspace
    grows by one row block's digit planes (6 bytes per Kfu entry) when the INT8 route is on, the offsets of the
    regions in front of the scratch area do not move, and the Python pool keys workspaces by route."""
    import ctypes
    from edrgp_b200 import _lib, ops
    lib = _lib.load()
    names = ops.FixedSweep.REGIONS
    assert ops.get_stats_mode() in ops.STATS_MODES
    before = ops.get_stats_mode()
    try:
        ops.set_stats_mode('fp64')
        off64 = (ctypes.c_int64 * len(names))()
        n64 = lib.edrgp_fixed_layout(2000000, 64, 512, 524288, 1, off64)
        ops.set_stats_mode('int8x6')
        assert ops.get_stats_mode() == 'int8x6' and lib.edrgp_get_stats_mode() == 1
        off8 = (ctypes.c_int64 * len(names))()
        n8 = lib.edrgp_fixed_layout(2000000, 64, 512, 524288, 1, off8)
        planes = 524288 * 512 * 6
        assert n8 - n64 >= planes - (n64 - 8 * dict(zip(names, off64))['scratch']) and n8 > n64
        assert n8 >= lib.edrgp_inducing_stats_i8_workspace_bytes(524288, 512) >= planes
        i = names.index('scratch')
        assert list(off8)[:i + 1] == list(off64)[:i + 1]
        # m beyond the INT8 kernels: the FP64 route's layout
        big8 = lib.edrgp_fixed_layout(100000, 64, 4096, 65536, 1, off8)
        ops.set_stats_mode('fp64')
        assert big8 == lib.edrgp_fixed_layout(100000, 64, 4096, 65536, 1, off64)
        assert lib.edrgp_inducing_stats_i8_workspace_bytes(1000, 4096) == 0
        with pytest.raises(ValueError):
            ops.set_stats_mode('int8x5')
        assert lib.edrgp_set_stats_mode(7) != 0
    finally:
        ops.set_stats_mode(before)


def test_peer_exchange_layout_and_argument_checks():
    """The NVLink exchange's host-side entry points without a GPU: the buffer layout (two copies of every payload behind
    the flag words), rejected arguments, and that a single process never builds an exchange."""
    import ctypes
    from edrgp_b200 import _lib, dist
    lib = _lib.load()
    off = (ctypes.c_int64 * 4)()
    m, d, world = 512, 64, 8
    nbytes = lib.edrgp_peer_layout(m, d, world, off)
    flags, table, stats, gram = [int(o) for o in off]
    assert flags == 0 and table == 4 * 16 // 2                       # 4 rows of 16 int32 flag words
    assert stats - table == 2 * (4 * world)
    per_copy = m * m + m + 1
    assert gram - stats == 2 * (per_copy + (per_copy & 1))
    assert nbytes == 8 * (gram + 2 * d * d)
    assert lib.edrgp_peer_layout(m, d, 17, off) == 0                  # at most 16 ranks
    assert lib.edrgp_peer_layout(0, d, 2, off) == 0
    assert lib.edrgp_peer_alloc(0, None, None) != 0
    assert lib.edrgp_peer_open(None, None) != 0
    assert lib.edrgp_fixed_bind_peers(None, None, 0, 2, m, d) != 0
    assert lib.edrgp_peer_close(None) == 0 and lib.edrgp_peer_free(None) == 0
    assert dist.peer_exchange(m, d) is None                          # not a distributed job
