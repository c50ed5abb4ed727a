"""ctypes binding of libedrgp_b200.so (C ABI declared in include/edrgp_b200.h).

There is NO CPU fallback: if the library is missing, or a call fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libedrgp_b200.so')

_c_dp = ctypes.c_void_p      # device pointers travel as integers
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/edrgp_b200.h one to one
SIGNATURES = {
    'edrgp_version': (_int, []),
    'edrgp_last_error': (ctypes.c_char_p, []),
    'edrgp_sm_count': (_int, []),
    'edrgp_launch_count': (ctypes.c_uint64, []),
    'edrgp_fp64_probe': (_int, [_c_dp, _int, ctypes.POINTER(ctypes.c_double), _c_dp]),
    'edrgp_pack_bytes': (_sz, [_int, _int]),
    'edrgp_pack_inducing': (_int, [_c_dp, _c_dp, _c_dp, _dbl, _c_dp, _int, _int, _c_dp, _c_dp]),
    'edrgp_kuf': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _int, _dbl, _c_dp, _i64, _int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_pack_tf32_bytes': (_sz, [_int, _int]),
    'edrgp_pack_inducing_tf32': (_int, [_c_dp, _c_dp, _int, _int, _c_dp, _c_dp]),
    'edrgp_kuf_tf32x3': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _c_dp, _int, _dbl, _c_dp, _i64, _c_dp]),
    'edrgp_pack_grad_tf32_bytes': (_sz, [_int, _int]),
    'edrgp_pack_grad_tf32': (_int, [_c_dp, _c_dp, _c_dp, _dbl, _int, _int, _c_dp, _c_dp]),
    'edrgp_grad_tf32x3': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _i64, _dbl, _c_dp, _c_dp, _int, _c_dp, _i64, _c_dp]),
    'edrgp_pack_weights_tf32_bytes': (_sz, [_int]),
    'edrgp_pack_weights_tf32': (_int, [_c_dp, _i64, _dbl, _int, _c_dp, _c_dp]),
    'edrgp_weights_tf32x3': (_int, [_c_dp, _i64, _int, _i64, _c_dp, _c_dp, _c_dp, _dbl, _c_dp, _i64, _c_dp, _c_dp]),
    'edrgp_grad_gram_workspace_bytes': (_sz, [_int]),
    'edrgp_grad_gram': (_int, [_c_dp, _i64, _int, _c_dp, _int, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_grad_gram_cached': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _i64, _dbl, _c_dp, _int, _c_dp, _i64, _c_dp,
                                      _c_dp, _c_dp]),
    'edrgp_syrk_workspace_bytes': (_sz, [_i64, _int]),
    'edrgp_syrk': (_int, [_c_dp, _i64, _int, _i64, _c_dp, _i64, _int, _c_dp, _c_dp]),
    'edrgp_inducing_stats': (_int, [_c_dp, _i64, _int, _i64, _c_dp, _c_dp, _i64, _c_dp, _int, _c_dp, _c_dp]),
    'edrgp_inducing_stats_i8_workspace_bytes': (_sz, [_i64, _int]),
    'edrgp_inducing_stats_i8': (_int, [_c_dp, _i64, _int, _i64, _c_dp, _dbl, _c_dp, _i64, _c_dp, _int, _c_dp, _c_dp]),
    'edrgp_gemm_tn_workspace_bytes': (_sz, [_i64, _int, _int]),
    'edrgp_gemm_tn': (_int, [_c_dp, _i64, _int, _c_dp, _i64, _int, _i64, _c_dp, _i64, _int, _c_dp, _c_dp]),
    'edrgp_kmm': (_int, [_c_dp, _i64, _c_dp, _int, _int, _dbl, _dbl, _c_dp, _i64, _int, _int, _c_dp]),
    'edrgp_solve_workspace_bytes': (_sz, [_int]),
    'edrgp_solve': (_int, [_c_dp, _c_dp, _c_dp, _int, _dbl, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_vfe_grad_small_workspace_bytes': (_sz, [_int]),
    'edrgp_vfe_grad_small': (_int, [_c_dp, _c_dp, _c_dp, _c_dp, _int, _dbl, _c_dp, _i64, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_potrf': (_int, [_c_dp, _int, _i64, _c_dp, _c_dp]),
    'edrgp_posv': (_int, [_c_dp, _int, _i64, _c_dp, _i64, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_trsm': (_int, [_c_dp, _int, _c_dp, _int, _int, _c_dp]),
    'edrgp_eigh_workspace_bytes': (_sz, [_int]),
    'edrgp_eigh': (_int, [_c_dp, _int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_count_nonfinite': (_int, [_c_dp, _i64, _c_dp, _c_dp]),
    'edrgp_col_moments_workspace_bytes': (_sz, [_int]),
    'edrgp_col_moments': (_int, [_c_dp, _i64, _int, _c_dp, _c_dp, _c_dp, _int, _c_dp, _c_dp]),
    'edrgp_weights_workspace_bytes': (_sz, [_i64, _int]),
    'edrgp_weights': (_int, [_c_dp, _i64, _int, _i64, _c_dp, _i64, _c_dp, _c_dp, _dbl, _dbl, _c_dp, _i64, _c_dp,
                             _c_dp, _int, _c_dp, _c_dp]),
    'edrgp_standardize': (_int, [_c_dp, _i64, _int, _c_dp, _c_dp, _c_dp, _c_dp]),
    'edrgp_project_dmma': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _int, _c_dp, _i64, _c_dp]),
    'edrgp_project': (_int, [_c_dp, _i64, _int, _c_dp, _int, _c_dp, _c_dp]),
    'edrgp_fixed_layout': (_sz, [_i64, _int, _int, _i64, _int, ctypes.POINTER(ctypes.c_int64)]),
    'edrgp_set_stats_mode': (_int, [_int]),
    'edrgp_get_stats_mode': (_int, []),
    'edrgp_fixed_begin': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _c_dp, _i64, _c_dp, _int, _dbl, _i64, _c_dp, _i64, _int,
                                 _int, _c_dp, _i64, _c_dp, _c_dp]),
    'edrgp_fixed_stats': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _int, _dbl, _i64, _c_dp, _i64, _int, _int, _c_dp, _i64,
                                 _c_dp, _c_dp]),
    'edrgp_fixed_posterior': (_int, [_c_dp, _i64, _i64, _int, _int, _dbl, _dbl, _dbl, _i64, _int, _c_dp, _c_dp]),
    'edrgp_fixed_grad': (_int, [_c_dp, _i64, _i64, _int, _c_dp, _i64, _c_dp, _i64, _c_dp, _int, _dbl, _dbl, _c_dp, _c_dp,
                                _i64, _i64, _int, _c_dp, _c_dp]),
    'edrgp_fixed_eigh': (_int, [_i64, _int, _int, _i64, _int, _c_dp, _c_dp]),
    'edrgp_fixed_reduce_gram': (_int, [_i64, _int, _int, _i64, _int, _c_dp, _c_dp]),
    'edrgp_peer_layout': (_sz, [_int, _int, _int, ctypes.POINTER(_i64)]),
    'edrgp_peer_alloc': (_int, [_sz, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    'edrgp_peer_open': (_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'edrgp_peer_close': (_int, [_c_dp]),
    'edrgp_peer_free': (_int, [_c_dp]),
    'edrgp_fixed_bind_peers': (_int, [_c_dp, ctypes.POINTER(ctypes.c_void_p), _int, _int, _int, _int]),
    'edrgp_h2d_open': (ctypes.c_void_p, [_c_dp, _c_dp, _i64, _sz, _sz, _i64, _int, _int, _c_dp, _c_dp, _c_dp, _sz]),
    'edrgp_h2d_wait_side': (_int, [_c_dp, _c_dp]),
    'edrgp_h2d_wait': (_int, [_c_dp, _i64, _i64, _c_dp]),
    'edrgp_h2d_staged': (_int, [_c_dp]),
    'edrgp_h2d_close': (_int, [_c_dp]),
    'edrgp_timing_begin': (_int, []),
    'edrgp_timing_end': (_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]),
}


class EdrgpError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EdrgpError(
            "libedrgp_b200.so not found at %s: build it with `python -m edrgp_b200.build` "
            "(there is no CPU fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().edrgp_last_error()
        raise EdrgpError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ''))
