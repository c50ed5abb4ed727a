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


# optional per-op CUDA-event timing (bench.py's roofline): name -> [(start, end), ...]
_TIMING = None


class _Timed(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _TIMING.setdefault(self.name, []).append((self.e0, e1))
        return False


_STAGES = ('kuf', 'targets', 'inducing_stats', 'solve', 'grad_gram_cached', 'eigh')      # EDRGP_STAGE_* order


def start_timing():
    """Begin collecting CUDA-event pairs around the contraction ops on the current stream (the composite
    calls of the fixed sweep bracket their stages themselves: edrgp_timing_begin)."""
    global _TIMING
    _TIMING = {}
    _lib.check(_lib.load().edrgp_timing_begin(), 'edrgp_timing_begin')


def stop_timing():
    """Stop collecting; returns {op name: (total ms, launches)} (synchronises)."""
    import ctypes
    global _TIMING
    t, _TIMING = _TIMING, None
    torch.cuda.synchronize()
    out = {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in (t or {}).items()}
    ms = (ctypes.c_double * len(_STAGES))()
    cnt = (ctypes.c_int * len(_STAGES))()
    _lib.check(_lib.load().edrgp_timing_end(ms, cnt), 'edrgp_timing_end')
    for name, a, c in zip(_STAGES, ms, cnt):
        if c:
            prev = out.get(name, (0.0, 0))
            out[name] = (prev[0] + a, prev[1] + c)
    return out


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
    """Device-resident tiled copy of (Z / l^2, -|z/l|^2 / 2, coef): see edrgp_pack_inducing.

    ``block`` features per kernel call (128 for the cross-covariance kernels, 64 for the cached
    gradient kernel): a wider Z is packed as several feature blocks, each with its own buffer;
    exp(-r^2/2) factorises over the blocks and the gradient of a feature only needs its own block.
    """

    def __init__(self, Z, ell, coef=None, coef_scale=1.0, block=128, dev_scale=None):
        _need_cuda(Z, ell, coef, dev_scale)
        self.m, d = Z.shape
        if d % 2:
            Z = pad_even(Z)
            ell = torch.cat([ell, torch.ones(1, dtype=F64, device=ell.device)])
        self.d = Z.shape[1]
        self.Z, self.ell = Z, ell
        self.block = int(block)
        self.blocks = []                      # (first feature, width, Z block, ell block, buffer)
        lib = _lib.load()
        for c0 in range(0, self.d, self.block):
            dc = min(self.block, self.d - c0)
            if self.d <= self.block:
                Zb, eb = Z, ell
            else:
                Zb, eb = Z[:, c0:c0 + dc].contiguous(), ell[c0:c0 + dc].contiguous()
            buf = torch.empty(lib.edrgp_pack_bytes(self.m, dc) // 8, dtype=F64, device=Z.device)
            self.blocks.append((c0, dc, Zb, eb, buf))
        self.buf = self.blocks[0][4]
        self.set_coef(coef, coef_scale, dev_scale)

    def set_coef(self, coef, coef_scale=1.0, dev_scale=None):
        """coef * coef_scale [* dev_scale[0], a 1-element device tensor: no read-back, no extra kernel]."""
        lib = _lib.load()
        _need_cuda(coef, dev_scale)
        for c0, dc, Zb, eb, buf in self.blocks:
            _lib.check(lib.edrgp_pack_inducing(_ptr(Zb), _ptr(eb), _ptr(coef), float(coef_scale), _ptr(dev_scale),
                                               self.m, dc, _ptr(buf), _stream()), 'edrgp_pack_inducing')
        return self


def kuf(X, pack, sf2, y=None, out=None, want_K=True, want_mu=False, flag=None):
    """Kfu (n, m) and, if y is given, b = Kfu^T y.  want_mu: returns (K, b, mu) with
    mu = Kfu @ coef (the coefficients stored in the pack).  Any feature count: more than 128
    features are evaluated block by block into the same buffer (multiply mode).  flag: 1-element int32
    device tensor set to 1 when a row of X holds a NaN / Inf (the scan of check_X_y for free)."""
    lib = _lib.load()
    X = pad_even(X)
    _need_cuda(X, y)
    n, ldx = X.shape
    if ldx != pack.d:
        raise ValueError("X has %d features, the pack %d" % (ldx, pack.d))
    K = None
    ldk = pack.m + (pack.m & 1)
    nblk = len(pack.blocks)
    if want_K or nblk > 1:
        K = out if out is not None else torch.empty(n, ldk, dtype=F64, device=X.device)
    b = torch.zeros(pack.m, dtype=F64, device=X.device) if y is not None else None
    mu = torch.empty(n, dtype=F64, device=X.device) if want_mu else None
    with _Timed('kuf'):
        for i, (c0, dc, _, _, buf) in enumerate(pack.blocks):
            last = i == nblk - 1
            _lib.check(lib.edrgp_kuf(X.data_ptr() + 8 * c0, ldx, n, dc, _ptr(buf), pack.m,
                                     float(sf2) if last else 1.0, _ptr(K), ldk, int(i > 0),
                                     _ptr(y) if last else 0, _ptr(b) if last else 0, _ptr(mu) if last else 0,
                                     _ptr(flag), _stream()), 'edrgp_kuf')
    if not want_K:
        K = None
    if K is not None and ldk != pack.m:
        K = K[:, :pack.m]
    if want_mu:
        return K, b, mu
    return K, b


class InducingPackTF32(object):
    """Device image of Z / l for the TF32-split cross-covariance kernel (edrgp_pack_inducing_tf32)."""

    def __init__(self, Z, ell):
        _need_cuda(Z, ell)
        lib = _lib.load()
        self.m, self.d = Z.shape
        if self.d % 2:
            Z = pad_even(Z)
            ell = torch.cat([ell, torch.ones(1, dtype=F64, device=ell.device)])
            self.d += 1
        if self.d > 64:
            raise ValueError("the TF32-split mode covers d <= 64 (got %d)" % self.d)
        self.ell = ell.contiguous()
        self.buf = torch.empty(lib.edrgp_pack_tf32_bytes(self.m, self.d) // 8, dtype=F64, device=Z.device)
        _lib.check(lib.edrgp_pack_inducing_tf32(_ptr(Z), _ptr(self.ell), self.m, self.d, _ptr(self.buf), _stream()),
                   'edrgp_pack_inducing_tf32')


def kuf_tf32(X, pack, sf2, out=None):
    """Kfu (n, m) in the TF32-split mode: tcgen05 distance contraction, FP64 entries out."""
    lib = _lib.load()
    X = pad_even(X)
    _need_cuda(X)
    n, ldx = X.shape
    if ldx != pack.d:
        raise ValueError("X has %d features, the pack %d" % (ldx, pack.d))
    ldk = pack.m + (pack.m & 1)
    K = out if out is not None else torch.empty(n, ldk, dtype=F64, device=X.device)
    with _Timed('kuf'):
        _lib.check(lib.edrgp_kuf_tf32x3(_ptr(X), ldx, n, ldx, _ptr(pack.ell), _ptr(pack.buf), pack.m, float(sf2),
                                        _ptr(K), K.shape[1], _stream()), 'edrgp_kuf_tf32x3')
    return K if ldk == pack.m else K[:, :pack.m]


def grad_tf32(X, K, Z, ell, coef, coef_scale, sf2, G_out=None):
    """Posterior-mean gradients G (n, d) from a stored cross-covariance block K (n, ldk) in the TF32-split
    mode: one tcgen05 contraction K x [c | c z / l^2] (c = coef * coef_scale), FP64 correction and output."""
    lib = _lib.load()
    d_user = X.shape[1]
    X = pad_even(X)
    Z = pad_even(Z)
    if ell.shape[0] != X.shape[1]:
        ell = torch.cat([ell, torch.ones(X.shape[1] - ell.shape[0], dtype=F64, device=ell.device)])
    _need_cuda(X, K, Z, ell, coef)
    n, d = X.shape
    m = Z.shape[0]
    if d > 64:
        raise ValueError("the TF32-split mode covers d <= 64 (got %d)" % d)
    if K.shape[0] != n or K.shape[1] < m:
        raise ValueError("K and X / Z shapes differ")
    pack = torch.empty(lib.edrgp_pack_grad_tf32_bytes(m, d) // 8, dtype=F64, device=X.device)
    _lib.check(lib.edrgp_pack_grad_tf32(_ptr(Z), _ptr(ell), _ptr(coef), float(coef_scale), m, d, _ptr(pack), _stream()),
               'edrgp_pack_grad_tf32')
    G = G_out if (G_out is not None and d == d_user) else torch.empty(n, d, dtype=F64, device=X.device)
    with _Timed('grad_tf32'):
        _lib.check(lib.edrgp_grad_tf32x3(_ptr(X), d, n, d, _ptr(K), K.shape[1], float(sf2), _ptr(ell), _ptr(pack), m,
                                         _ptr(G), G.shape[1], _stream()), 'edrgp_grad_tf32x3')
    return G if d == d_user else G[:, :d_user].contiguous()


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
    with _Timed('grad_gram'):
        _lib.check(lib.edrgp_grad_gram(_ptr(X), n, d, _ptr(pack.buf), pack.m, _ptr(G), _ptr(C), _ptr(ws),
                                       _stream()), 'edrgp_grad_gram')
    if d != d_user:
        if G is not None:
            G = G[:, :d_user].contiguous()
        if C is not None:
            C = C[:d_user, :d_user].contiguous()
    return G, C


def grad_gram_cached(X, K, pack, sf2, want_G=True, want_C=True, G_out=None):
    """``grad_gram`` from a stored cross-covariance block K (n, ldk) (entries sf2 exp(-r^2/2), as
    written by ``kuf``); ``pack`` carries coef = alpha * scale (without sf2) in feature blocks of at
    most 64.  One block: G and C come from the fused kernel.  Several blocks (d > 64): every block
    writes its columns of G and C = G^T G runs on the symmetric reduction."""
    lib = _lib.load()
    d_user = X.shape[1]
    X = pad_even(X)
    _need_cuda(X, K)
    n, d = X.shape
    if K.shape[0] != n:
        raise ValueError("K and X row counts differ")
    if d != pack.d or max(dc for _, dc, _, _, _ in pack.blocks) > 64:
        raise ValueError("the pack must hold the same features in blocks of at most 64")
    single = len(pack.blocks) == 1
    G = None
    if want_G or not single:
        G = G_out if (G_out is not None and d == d_user) else torch.empty(n, d, dtype=F64, device=X.device)
    C = ws = None
    if want_C and single:
        C = torch.empty(d, d, dtype=F64, device=X.device)
        ws = torch.empty(lib.edrgp_grad_gram_workspace_bytes(d) // 8, dtype=F64, device=X.device)
    with _Timed('grad_gram_cached'):
        for c0, dc, _, _, buf in pack.blocks:
            _lib.check(lib.edrgp_grad_gram_cached(X.data_ptr() + 8 * c0, d, n, dc, _ptr(K), K.shape[1], float(sf2),
                                                  _ptr(buf), pack.m, 0 if G is None else G.data_ptr() + 8 * c0, d,
                                                  _ptr(C), _ptr(ws), _stream()), 'edrgp_grad_gram_cached')
    if want_C and not single:
        C = syrk(G)
    if d != d_user:
        if G is not None:
            G = G[:, :d_user].contiguous()
        if C is not None:
            C = C[:d_user, :d_user].contiguous()
    return (G if want_G else None), C


def syrk(A, k=None, out=None, accumulate=False):
    """C (+)= A[:, :k]^T A[:, :k] for a tall (n, lda) tensor (lda even)."""
    lib = _lib.load()
    _need_cuda(A, out)
    n, lda = A.shape
    k = lda if k is None else k
    if lda % 2:
        raise ValueError("syrk needs an even leading dimension")
    C = out if out is not None else torch.empty(k, k, dtype=F64, device=A.device)
    ws = torch.empty(max(1, lib.edrgp_syrk_workspace_bytes(n, k) // 8), dtype=F64, device=A.device)
    _lib.check(lib.edrgp_syrk(_ptr(A), n, k, lda, _ptr(C), k, int(bool(accumulate and out is not None)),
                              _ptr(ws), _stream()), 'edrgp_syrk')
    return C


def inducing_stats(K, y, m=None, P=None, b_yy=None, accumulate=False):
    """P (+)= K^T K, b_yy[:m] (+)= K^T y, b_yy[m] (+)= y^T y from a stored cross-covariance block
    K (n, ldk) (ldk even, first m columns used)."""
    lib = _lib.load()
    _need_cuda(K, y, P, b_yy)
    n, ldk = K.shape
    m = ldk if m is None else m
    if ldk % 2:
        raise ValueError("inducing_stats needs an even leading dimension")
    fresh = P is None or b_yy is None
    if P is None:
        P = torch.empty(m, m, dtype=F64, device=K.device)
    if b_yy is None:
        b_yy = torch.empty(m + 1, dtype=F64, device=K.device)
    ws = torch.empty(max(1, lib.edrgp_syrk_workspace_bytes(n, m) // 8), dtype=F64, device=K.device)
    with _Timed('inducing_stats'):
        _lib.check(lib.edrgp_inducing_stats(_ptr(K), n, m, ldk, _ptr(y), _ptr(P), m, _ptr(b_yy),
                                            int(bool(accumulate and not fresh)), _ptr(ws), _stream()),
                   'edrgp_inducing_stats')
    return P, b_yy


def inducing_stats_i8(K, y, sf2, m=None, P=None, b_yy=None, accumulate=False):
    """``inducing_stats`` on the INT8 tensor cores (exact products of six signed radix-256 digits of K / (4 sf2)):
    K (n, ldk) must hold kernel entries in [0, sf2] as written by ``kuf`` with the same ``sf2``."""
    lib = _lib.load()
    _need_cuda(K, y, P, b_yy)
    n, ldk = K.shape
    m = ldk if m is None else m
    fresh = P is None or (y is not None and b_yy is None)
    if P is None:
        P = torch.empty(m, m, dtype=F64, device=K.device)
    if b_yy is None and y is not None:
        b_yy = torch.empty(m + 1, dtype=F64, device=K.device)
    nbytes = lib.edrgp_inducing_stats_i8_workspace_bytes(n, m)
    if nbytes == 0:
        raise ValueError("the INT8 statistics cover m <= 2048 (got %d)" % m)
    ws = torch.empty(nbytes // 8, dtype=F64, device=K.device)
    with _Timed('inducing_stats'):
        _lib.check(lib.edrgp_inducing_stats_i8(_ptr(K), n, m, ldk, _ptr(y), float(sf2), _ptr(P), m, _ptr(b_yy),
                                               int(bool(accumulate and not fresh)), _ptr(ws), _stream()),
                   'edrgp_inducing_stats_i8')
    return P, b_yy


def gemm_tn(A, B, ka=None, kb=None, out=None, accumulate=False):
    """C (+)= A[:, :ka]^T B[:, :kb] for tall row-major A (n, lda), B (n, ldb)."""
    lib = _lib.load()
    _need_cuda(A, B, out)
    n, lda = A.shape
    ldb = B.shape[1]
    ka = lda if ka is None else ka
    kb = ldb if kb is None else kb
    if lda % 2 or ldb % 2 or B.shape[0] != n:
        raise ValueError("gemm_tn needs even leading dimensions and matching row counts")
    C = out if out is not None else torch.empty(ka, kb, dtype=F64, device=A.device)
    ws = torch.empty(max(1, lib.edrgp_gemm_tn_workspace_bytes(n, ka, kb) // 8), dtype=F64, device=A.device)
    with _Timed('gemm_tn'):
        _lib.check(lib.edrgp_gemm_tn(_ptr(A), lda, ka, _ptr(B), ldb, kb, n, _ptr(C), kb,
                                     int(bool(accumulate and out is not None)), _ptr(ws), _stream()), 'edrgp_gemm_tn')
    return C


def kmm(pack, sf2, jitter=1e-8):
    """Kuu = K(Z, Z) + jitter I with GPy's exact-sf2 diagonal."""
    lib = _lib.load()
    m = pack.m
    ldk = m + (m & 1)
    K = torch.empty(m, ldk, dtype=F64, device=pack.buf.device)
    nblk = len(pack.blocks)
    for i, (c0, dc, _, _, buf) in enumerate(pack.blocks):
        # the variance rides on the last block: its call also finishes the diagonal with sf2 + jitter
        _lib.check(lib.edrgp_kmm(pack.Z.data_ptr() + 8 * c0, pack.d, _ptr(buf), m, dc,
                                 float(sf2) if i == nblk - 1 else 1.0, float(jitter), _ptr(K), ldk, int(i > 0),
                                 int(i == nblk - 1), _stream()), 'edrgp_kmm')
    return K if ldk == m else K[:, :m].contiguous()


class SolveResult(object):
    __slots__ = ('alpha', 'c', 'Lm', 'LB', 'B', 'scalars', 'info')


def solve(Kmm, P, b, beta):
    """VarDTC solve chain on the device.  Kmm is consumed (overwritten by its Cholesky factor)."""
    lib = _lib.load()
    _need_cuda(Kmm, P, b)
    m = Kmm.shape[0]
    dev = Kmm.device
    out = SolveResult()
    out.Lm = Kmm
    out.LB = torch.empty(m, m, dtype=F64, device=dev)
    out.alpha = torch.empty(m, dtype=F64, device=dev)
    out.c = torch.empty(m, dtype=F64, device=dev)
    out.scalars = torch.empty(4, dtype=F64, device=dev)
    out.info = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.edrgp_solve_workspace_bytes(m) // 8, dtype=F64, device=dev)
    out.B = ws[:m * m].view(m, m)            # I + A = I + beta Lm^-1 P Lm^-T, kept by the chain
    with _Timed('solve'):
        _lib.check(lib.edrgp_solve(_ptr(Kmm), _ptr(P), _ptr(b), m, float(beta), _ptr(out.LB), _ptr(out.alpha),
                                   _ptr(out.c), _ptr(out.scalars), _ptr(out.info), _ptr(ws), _stream()), 'edrgp_solve')
    return out


def vfe_grad_small(res, beta):
    """The m x m chain of the bound's gradients from a ``solve`` result: (Msym with an even leading dimension,
    Dsym, sumAE as a device scalar) -- see edrgp_vfe_grad_small."""
    lib = _lib.load()
    m = res.Lm.shape[0]
    dev = res.Lm.device
    ldm = m + (m & 1)
    Msym = torch.zeros(m, ldm, dtype=F64, device=dev) if ldm != m else torch.empty(m, m, dtype=F64, device=dev)
    Dsym = torch.empty(m, m, dtype=F64, device=dev)
    sumAE = torch.empty(1, dtype=F64, device=dev)
    ws = torch.empty(lib.edrgp_vfe_grad_small_workspace_bytes(m) // 8, dtype=F64, device=dev)
    with _Timed('solve'):
        _lib.check(lib.edrgp_vfe_grad_small(_ptr(res.LB), _ptr(res.Lm), _ptr(res.B), _ptr(res.c), m, float(beta),
                                            _ptr(Msym), ldm, _ptr(Dsym), _ptr(sumAE), _ptr(ws), _stream()),
                   'edrgp_vfe_grad_small')
    return Msym, Dsym, sumAE


def potrf(A):
    """In-place lower Cholesky of a symmetric (m, m) tensor; returns (A, info) with info a device int32
    (0, or 1 + index of the first non-positive pivot)."""
    lib = _lib.load()
    _need_cuda(A)
    m = A.shape[0]
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    with _Timed('solve'):
        _lib.check(lib.edrgp_potrf(_ptr(A), m, A.shape[1], _ptr(info), _stream()), 'edrgp_potrf')
    return A, info


def posv(A, rhs):
    """Solve A x = rhs (one right-hand side) by Cholesky: returns (x, L, info).  A's lower triangle and rhs
    are consumed; L is a new (m, m) tensor whose lower triangle is the factor."""
    lib = _lib.load()
    _need_cuda(A, rhs)
    m = A.shape[0]
    L = torch.empty(m, m, dtype=F64, device=A.device)
    x = torch.empty(m, dtype=F64, device=A.device)
    info = torch.empty(1, dtype=torch.int32, device=A.device)          # zeroed by the call
    with _Timed('solve'):
        _lib.check(lib.edrgp_posv(_ptr(A), m, A.shape[1], _ptr(L), m, _ptr(rhs), _ptr(x), _ptr(info), _stream()),
                   'edrgp_posv')
    return x, L, info


def trsm(L, B, trans=False):
    """In-place triangular solve with a lower factor: L X = B (trans=False) or L^T X = B."""
    lib = _lib.load()
    _need_cuda(L, B)
    m = L.shape[0]
    nrhs = 1 if B.dim() == 1 else B.shape[1]
    _lib.check(lib.edrgp_trsm(_ptr(L), m, _ptr(B), nrhs, int(bool(trans)), _stream()), 'edrgp_trsm')
    return B


EIGH_MAX_SWEEPS = 60


def eigh(C, return_status=False):
    """Descending eigen-decomposition of a symmetric (d, d) matrix: (evals, comps) with comps rows
    = eigenvectors.  C is not modified.  return_status: also the device int32 pair [sweeps, barrier time-out]
    for ``check_eigh_status`` once it has reached the host (callers read it with the eigenvalues)."""
    lib = _lib.load()
    _need_cuda(C)
    d = C.shape[0]
    A = C.contiguous()
    evals = torch.empty(d, dtype=F64, device=C.device)
    comps = torch.empty(d, d, dtype=F64, device=C.device)
    ws = torch.empty(max(1, lib.edrgp_eigh_workspace_bytes(d) // 8), dtype=F64, device=C.device)
    status = torch.zeros(2, dtype=torch.int32, device=C.device) if return_status else None
    with _Timed('eigh'):
        _lib.check(lib.edrgp_eigh(_ptr(A), d, _ptr(evals), _ptr(comps), _ptr(status), _ptr(ws), _stream()), 'edrgp_eigh')
    return (evals, comps, status) if return_status else (evals, comps)


def check_eigh_status(sweeps, timed_out):
    """Raise instead of handing out half-rotated eigenvectors."""
    if int(timed_out) != 0:
        raise _lib.EdrgpError("edrgp_eigh: the multi-SM Jacobi solver's grid barrier timed out (a CTA was not resident)")
    if int(sweeps) >= EIGH_MAX_SWEEPS:
        raise _lib.EdrgpError("edrgp_eigh: no convergence within %d Jacobi sweeps" % EIGH_MAX_SWEEPS)


def col_moments(X, shift=None, weight=None, out=None, accumulate=False):
    """(sum_i w_i (x - shift), sum_i w_i (x - shift)^2) per column, each of shape (d,)."""
    lib = _lib.load()
    if X.dim() == 1:
        X = X[:, None]
    _need_cuda(X, shift, weight, out)
    n, d = X.shape
    acc = int(bool(accumulate and out is not None))
    if out is None:
        out = torch.empty(2 * d, dtype=F64, device=X.device)
    ws = torch.empty(lib.edrgp_col_moments_workspace_bytes(d) // 8, dtype=F64, device=X.device)
    with _Timed('col_moments'):
        _lib.check(lib.edrgp_col_moments(_ptr(X), n, d, _ptr(shift), _ptr(weight), _ptr(out), acc, _ptr(ws),
                                         _stream()), 'edrgp_col_moments')
    return out[:d], out[d:]


def even_ld(M):
    """Row-major (r, c) tensor with an even leading dimension (zero padding column if c is odd)."""
    if M.shape[1] % 2 == 0:
        return M.contiguous()
    out = torch.zeros(M.shape[0], M.shape[1] + 1, dtype=F64, device=M.device)
    out[:, :M.shape[1]] = M
    return out


def weights(K, M, m=None, y=None, alpha=None, c_ya=0.0, c_km=1.0, T=None, want_rowsum=False, colsum=None,
            accumulate=False):
    """T = K o (c_ya y alpha^T + c_km K M); returns (rowsum or None).  K (n, ldk), M (m, ldm) with even
    leading dimensions; T (n, ldt) is filled when given; colsum (m) is (+)= T^T 1 when given."""
    lib = _lib.load()
    _need_cuda(K, M, y, alpha, T, colsum)
    n, ldk = K.shape
    m = ldk if m is None else m
    rowsum = torch.empty(n, dtype=F64, device=K.device) if want_rowsum else None
    ws = None
    if want_rowsum or colsum is not None:
        ws = torch.empty(max(1, lib.edrgp_weights_workspace_bytes(n, m) // 8), dtype=F64, device=K.device)
    with _Timed('weights'):
        _lib.check(lib.edrgp_weights(_ptr(K), n, m, ldk, _ptr(M), M.shape[1], _ptr(y), _ptr(alpha), float(c_ya),
                                     float(c_km), _ptr(T), 0 if T is None else T.shape[1], _ptr(rowsum), _ptr(colsum),
                                     int(bool(accumulate)), _ptr(ws), _stream()), 'edrgp_weights')
    return rowsum


def weights_tf32(K, M, m=None, y=None, alpha=None, c_ya=0.0, c_km=1.0, T=None, want_rowsum=False):
    """``weights`` in the TF32-split mode (m <= 512): T = K o (c_ya y alpha^T + c_km K M) with the K M
    contraction on tcgen05; returns rowsum or None.  Column sums: ``col_moments(T)``."""
    lib = _lib.load()
    _need_cuda(K, M, y, alpha, T)
    n, ldk = K.shape
    m = ldk if m is None else m
    pack = torch.empty(lib.edrgp_pack_weights_tf32_bytes(m) // 8, dtype=F64, device=K.device)
    if pack.numel() == 0:
        raise ValueError("the TF32-split weights cover m <= 512 (got %d)" % m)
    _lib.check(lib.edrgp_pack_weights_tf32(_ptr(M), M.shape[1], float(c_km), m, _ptr(pack), _stream()),
               'edrgp_pack_weights_tf32')
    rowsum = torch.empty(n, dtype=F64, device=K.device) if want_rowsum else None
    if T is None:
        T = torch.empty(n, ldk, dtype=F64, device=K.device)
    with _Timed('weights'):
        _lib.check(lib.edrgp_weights_tf32x3(_ptr(K), n, m, ldk, _ptr(pack), _ptr(y), _ptr(alpha), float(c_ya), _ptr(T),
                                            0 if T is None else T.shape[1], _ptr(rowsum), _stream()),
                   'edrgp_weights_tf32x3')
    return rowsum


def count_nonfinite(*tensors):
    """Device int32 tensor holding the number of NaN / Inf entries over all given tensors."""
    lib = _lib.load()
    count = None
    for t in tensors:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == F64):
            raise ValueError("expected CUDA float64 tensors")
        t = t if t.is_contiguous() else t.contiguous()
        if count is None:
            count = torch.zeros(1, dtype=torch.int32, device=t.device)
        _lib.check(lib.edrgp_count_nonfinite(_ptr(t), t.numel(), _ptr(count), _stream()), 'edrgp_count_nonfinite')
    return count


def standardize(X, mean, scale, out=None):
    lib = _lib.load()
    X2 = X[:, None] if X.dim() == 1 else X
    _need_cuda(X2, mean, scale)
    n, d = X2.shape
    res = torch.empty_like(X2) if out is None else out
    _lib.check(lib.edrgp_standardize(_ptr(X2), n, d, _ptr(mean), _ptr(scale), _ptr(res), _stream()),
               'edrgp_standardize')
    return res[:, 0] if X.dim() == 1 else res


def project(X, V):
    """X (n, d) @ V^T with V (k, d): EDR.transform.  Few components: one streaming pass (a warp per
    row); many components (k > 8, d <= 128): the FP64 tensor-pipe contraction of the Kfu kernel."""
    lib = _lib.load()
    _need_cuda(X, V)
    n, d = X.shape
    k = V.shape[0]
    if k > 8 and d + (d & 1) <= 128 and n > 0:
        Xe = pad_even(X)
        pack = InducingPack(V.contiguous(), torch.ones(d, dtype=F64, device=X.device))
        ldo = k + (k & 1)
        out = torch.empty(n, ldo, dtype=F64, device=X.device)
        with _Timed('project'):
            _lib.check(lib.edrgp_project_dmma(_ptr(Xe), Xe.shape[1], n, Xe.shape[1], _ptr(pack.buf), k, _ptr(out), ldo,
                                              _stream()), 'edrgp_project_dmma')
        return out if ldo == k else out[:, :k].contiguous()
    out = torch.empty(n, k, dtype=F64, device=X.device)
    with _Timed('project'):
        _lib.check(lib.edrgp_project(_ptr(X), n, d, _ptr(V), k, _ptr(out), _stream()), 'edrgp_project')
    return out


STATS_MODES = ('fp64', 'int8x6')


def set_stats_mode(mode):
    """Statistics route of the composite sweep (edrgp_set_stats_mode): 'fp64' -- the FP64 DMMA reduction, the
    default -- or 'int8x6' -- exact integer products of six signed 8-bit digits on the INT8 tensor cores, P at FP64
    rounding level and ~1.5 x faster.  Process-wide; workspaces are pooled per mode.  The environment
    variable EDRGP_STATS sets the initial value."""
    if mode not in STATS_MODES:
        raise ValueError("stats mode must be one of %s (got %r)" % (STATS_MODES, mode))
    _lib.check(_lib.load().edrgp_set_stats_mode(STATS_MODES.index(mode)), 'edrgp_set_stats_mode')


def get_stats_mode():
    return STATS_MODES[_lib.load().edrgp_get_stats_mode()]


class FixedSweep(object):
    """The composite calls of the fixed-hyper-parameter sweep (``edrgp_fixed_*``) over ONE workspace tensor.

    Everything between two collectives is one C call; the regions a multi-rank caller all-reduces in place
    (``table``, ``stats``, ``C``) and the results (``alpha``, ``tail``, ``result``) are views of the workspace."""

    REGIONS = ('pack_k', 'pack_g', 'yt', 'stats', 'table', 'S', 'L', 'rhs', 'alpha', 'scratch', 'tail', 'result')

    _POOL = {}            # shape key -> workspaces of models that no longer exist, ready for reuse
    _POOL_DEPTH = 2

    @staticmethod
    def supported(d_even, n_local):
        return d_even % 2 == 0 and d_even <= 64 and n_local > 0

    @classmethod
    def acquire(cls, n_local, d, m, chunk_rows, rank, world, device):
        """A workspace for one model: a recycled one of the same shape when a previous model has been dropped
        (building the views costs more host time than the kernels of a small shard leave room for)."""
        key = (torch.device(device).index, int(n_local), int(d), int(m), int(chunk_rows), int(rank), int(world),
               get_stats_mode())
        free = cls._POOL.get(key)
        if free:
            fs = free.pop()
            fs.host = None
            return fs
        fs = cls(n_local, d, m, chunk_rows, rank, world, device)
        fs.key = key
        return fs

    @classmethod
    def release(cls, fs):
        """Called when the owning model is collected (weakref.finalize): nothing the model handed out aliases the
        workspace (``gradient_gram`` returns a copy of C), so it can serve the next model of the same shape."""
        free = cls._POOL.setdefault(fs.key, [])
        if len(free) < cls._POOL_DEPTH:
            free.append(fs)

    def __init__(self, n_local, d, m, chunk_rows, rank, world, device):
        import ctypes
        lib = _lib.load()
        self.n, self.d, self.m, self.chunk, self.rank, self.world = int(n_local), int(d), int(m), int(chunk_rows), rank, world
        off = (ctypes.c_int64 * len(self.REGIONS))()
        nbytes = lib.edrgp_fixed_layout(self.n, self.d, self.m, self.chunk, self.world, off)
        if nbytes == 0:
            raise _lib.EdrgpError("edrgp_fixed_layout rejected the shape")
        self.ws = torch.empty(nbytes // 8, dtype=F64, device=device)
        self.off = dict(zip(self.REGIONS, [int(o) for o in off]))
        m, d = self.m, self.d
        o = self.off
        self.table = self.ws[o['table']:o['table'] + 4 * world]
        self.stats = self.ws[o['stats']:o['stats'] + m * m + m + 1]
        self.P = self.stats[:m * m].view(m, m)
        self.byy = self.stats[m * m:]
        self.yt = self.ws[o['yt']:o['yt'] + self.n]
        self.alpha = self.ws[o['alpha']:o['alpha'] + m]
        self.L = self.ws[o['L']:o['L'] + m * m].view(m, m)
        self.tail = self.ws[o['tail']:o['tail'] + 4]
        self.result = self.ws[o['result']:o['result'] + d + 2 * d * d + 4]
        self.C = self.result[d + d * d:d + 2 * d * d].view(d, d)
        self.host = None                     # the result block once it has been read back (one transfer)
        self.key = None
        self.peer = None                     # dist.PeerExchange once bound (bind_peers): no collectives between the calls

    def bind_peers(self, exchange):
        """Bind this workspace to the job's NVLink exchange buffers (``dist.peer_exchange``): begin / stats_pass /
        posterior then reduce the moments table and the statistics over the ranks themselves and ``reduce_gram``
        replaces the all-reduce of C.  ``None`` unbinds."""
        lib = _lib.load()
        if exchange is None:
            _lib.check(lib.edrgp_fixed_bind_peers(_ptr(self.ws), None, 0, 0, 0, 0), 'edrgp_fixed_bind_peers')
        else:
            _lib.check(lib.edrgp_fixed_bind_peers(_ptr(self.ws), exchange.bases, self.rank, self.world, self.m, self.d),
                       'edrgp_fixed_bind_peers')
            if getattr(self, '_unbind', None) is None:
                # the binding is keyed by the workspace address: it must not outlive the tensor that owns it
                import weakref
                self._unbind = weakref.finalize(self, lib.edrgp_fixed_bind_peers, _ptr(self.ws), None, 0, 0, 0, 0)
        self.peer = exchange

    def reduce_gram(self):
        """C (result block) <- its sum over the ranks, read from the peers' buffers (bound workspaces only)."""
        lib = _lib.load()
        _lib.check(lib.edrgp_fixed_reduce_gram(self.n, self.d, self.m, *self._common()), 'edrgp_fixed_reduce_gram')
        self.host = None
        return self.C

    def _common(self):
        return self.chunk, self.world, _ptr(self.ws), _stream()

    def begin(self, X, y, Z, ell, sf2, Kfu, h2d=None, h2d_ahead=0):
        """h2d: an edrgp_h2d_open handle when X / y are still arriving from the host (the calls wait block by block)."""
        lib = _lib.load()
        _need_cuda(X, y, Z, ell, Kfu)
        _lib.check(lib.edrgp_fixed_begin(_ptr(X), X.shape[1], self.n, self.d, _ptr(y), _ptr(Z), Z.shape[1], _ptr(ell),
                                         self.m, float(sf2), self.chunk, _ptr(Kfu), Kfu.shape[1], self.rank, self.world,
                                         h2d, int(h2d_ahead), _ptr(self.ws), _stream()), 'edrgp_fixed_begin')
        self.host = None

    def stats_pass(self, X, y, sf2, Kfu, normalize, h2d=None, h2d_ahead=0):
        lib = _lib.load()
        _lib.check(lib.edrgp_fixed_stats(_ptr(X), X.shape[1], self.n, self.d, _ptr(y), self.m, float(sf2), self.chunk,
                                         _ptr(Kfu), Kfu.shape[1], int(bool(normalize)), self.world, h2d, int(h2d_ahead),
                                         _ptr(self.ws), _stream()), 'edrgp_fixed_stats')

    def posterior(self, Z, sf2, jitter, beta):
        lib = _lib.load()
        _lib.check(lib.edrgp_fixed_posterior(_ptr(Z), Z.shape[1], self.n, self.d, self.m, float(sf2), float(jitter),
                                             float(beta), *self._common()), 'edrgp_fixed_posterior')

    def grad(self, X, Kfu, Z, ell, sf2, coef_scale=1.0, dev_scale=None, G=None):
        """C = G^T G (view of the result block, NOT reduced over ranks) of the gradients with coefficients
        alpha * coef_scale [* dev_scale[0]]; G (n, ldg) is filled when given."""
        lib = _lib.load()
        _need_cuda(G, dev_scale)
        _lib.check(lib.edrgp_fixed_grad(_ptr(X), X.shape[1], self.n, self.d, _ptr(Kfu), Kfu.shape[1], _ptr(Z), Z.shape[1],
                                        _ptr(ell), self.m, float(sf2), float(coef_scale), _ptr(dev_scale), _ptr(G),
                                        0 if G is None else G.shape[1], *self._common()), 'edrgp_fixed_grad')
        self.host = None
        return self.C

    def eigh(self, C=None):
        """eigh of the (all-reduced) Gram matrix inside the result block, then ONE read-back of
        evals | components | C | tail.  C: the caller's copy of the matrix (possibly summed over ranks since
        ``grad`` returned it); it is put back into the block first."""
        lib = _lib.load()
        if C is not None and C.data_ptr() != self.C.data_ptr():
            self.C.copy_(C)
        _lib.check(lib.edrgp_fixed_eigh(self.n, self.d, self.m, *self._common()), 'edrgp_fixed_eigh')
        self.host = self.result.cpu().numpy()
        return self.host

    @staticmethod
    def decode_tail(tail4):
        """(non-finite flag, Cholesky info, N, mean, std) from the four tail doubles (numpy)."""
        import numpy as np
        ints = np.frombuffer(np.ascontiguousarray(tail4[:1]).tobytes(), dtype=np.int32)
        return int(ints[0]), int(ints[1]), float(tail4[1]), float(tail4[2]), float(tail4[3])


def launch_count():
    """Number of CUDA kernels the library has launched in this process."""
    return int(_lib.load().edrgp_launch_count())


def fp64_tensor_peak_tflops(iters=20000, reps=6):
    """Measured FP64 DMMA throughput (TFLOP/s) of the current device: CUDA-event timing of the
    register-only probe kernel, best of ``reps`` after one warm-up."""
    import ctypes
    lib = _lib.load()
    scratch = torch.zeros(256, dtype=F64, device='cuda')
    flops = ctypes.c_double(0.0)
    best = 0.0
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.edrgp_fp64_probe(_ptr(scratch), int(iters), ctypes.byref(flops), _stream()), 'edrgp_fp64_probe')
        e1.record()
        e1.synchronize()
        if r:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
