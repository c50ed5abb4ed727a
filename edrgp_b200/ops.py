"""Thin Python wrappers over the C ABI: torch owns device memory and streams, nothing else.

Every function takes/returns CUDA float64 torch tensors and enqueues on the current torch stream.
"""
import torch

from . import _lib

F64 = torch.float64


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == F64 and t.is_contiguous()):
            raise ValueError("expected contiguous CUDA float64 tensors")


def pad_even(X):
    """The kernels stream rows with 16-byte TMA bulk copies: odd d gets one zero column."""
    if X.shape[1] % 2 == 0:
        return X
    return torch.cat([X, torch.zeros(X.shape[0], 1, dtype=F64, device=X.device)], 1).contiguous()


class InducingPack(object):
    """Device-resident tiled copy of (Z / l^2, -|z/l|^2 / 2, coef): see edrgp_pack_inducing."""

    def __init__(self, Z, ell, coef=None, coef_scale=1.0):
        lib = _lib.load()
        _need_cuda(Z, ell, coef)
        self.m, d = Z.shape
        if d % 2:
            Z = pad_even(Z)
            ell = torch.cat([ell, torch.ones(1, dtype=F64, device=ell.device)])
        self.d = Z.shape[1]
        self.Z, self.ell = Z, ell
        nbytes = lib.edrgp_pack_bytes(self.m, self.d)
        self.buf = torch.empty(nbytes // 8, dtype=F64, device=Z.device)
        self.set_coef(coef, coef_scale)

    def set_coef(self, coef, coef_scale=1.0):
        lib = _lib.load()
        _need_cuda(coef)
        _lib.check(lib.edrgp_pack_inducing(_ptr(self.Z), _ptr(self.ell), _ptr(coef), float(coef_scale),
                                           self.m, self.d, _ptr(self.buf), _stream()), 'edrgp_pack_inducing')
        return self


def kuf(X, pack, sf2, y=None, out=None, want_K=True):
    """Kfu (n, m) and, if y is given, b = Kfu^T y."""
    lib = _lib.load()
    X = pad_even(X)
    _need_cuda(X, y)
    n = X.shape[0]
    K = None
    ldk = pack.m + (pack.m & 1)
    if want_K:
        K = out if out is not None else torch.empty(n, ldk, dtype=F64, device=X.device)
    b = torch.zeros(pack.m, dtype=F64, device=X.device) if y is not None else None
    _lib.check(lib.edrgp_kuf(_ptr(X), n, X.shape[1], _ptr(pack.buf), pack.m, float(sf2), _ptr(K), ldk,
                             _ptr(y), _ptr(b), _stream()), 'edrgp_kuf')
    if K is not None and ldk != pack.m:
        K = K[:, :pack.m]
    return K, b


def grad_gram(X, pack, want_G=True, want_C=True, G_out=None):
    """Posterior-mean gradients G (n, d) and/or their Gram matrix C = G^T G (d, d)."""
    lib = _lib.load()
    d_user = X.shape[1]
    X = pad_even(X)
    _need_cuda(X)
    n, d = X.shape
    G = None
    if want_G:
        G = G_out if (G_out is not None and d == d_user) else torch.empty(n, d, dtype=F64, device=X.device)
    C = ws = None
    if want_C:
        C = torch.empty(d, d, dtype=F64, device=X.device)
        ws = torch.empty(lib.edrgp_grad_gram_workspace_bytes(d) // 8, dtype=F64, device=X.device)
    _lib.check(lib.edrgp_grad_gram(_ptr(X), n, d, _ptr(pack.buf), pack.m, _ptr(G), _ptr(C), _ptr(ws),
                                   _stream()), 'edrgp_grad_gram')
    if d != d_user:
        if G is not None:
            G = G[:, :d_user].contiguous()
        if C is not None:
            C = C[:d_user, :d_user].contiguous()
    return G, C
