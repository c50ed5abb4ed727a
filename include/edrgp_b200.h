/*
 * edrgp_b200 -- C ABI of the B200-native sparse-GP posterior-gradient EDR hot path.
 *
 * The reference (neuro-ml/edr-gp) has no FFI: its boundary for this path is Python duck typing
 * (estimator.fit / predict_gradient, transformer.fit / components_; edrgp/base.py:133-134,
 * 163-167, 274, 334-343) and all arithmetic sits in GPy.  Each entry point below replaces one
 * arithmetic unit that the reference reaches through those calls; the comment on each names the
 * reference call site (file:line under the reference tree) and the GPy routine it stands for.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers to row-major FP64 buffers owned by the caller (torch tensors in
 *    the Python host), sizes as integers, `stream` is a cudaStream_t passed as void*.
 *  - every function only ENQUEUES work on `stream`: no allocation, no host synchronisation.
 *    Scratch comes from a caller-provided workspace whose size the *_workspace_bytes functions give.
 *  - return value 0 on success, negative on error; edrgp_last_error() returns a thread-local
 *    message for the last failing call.
 *  - re-entrant per (device, stream).
 */
#ifndef EDRGP_B200_H
#define EDRGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDRGP_OK 0
#define EDRGP_ERR_ARG (-1)
#define EDRGP_ERR_CUDA (-2)
#define EDRGP_ERR_UNSUPPORTED (-3)

/* precision modes of the contraction kernels */
#define EDRGP_FP64 0

int edrgp_version(void);
const char* edrgp_last_error(void);
/* number of SMs of the current device (grid sizing is done inside the library) */
int edrgp_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Inducing-point pack.  Replaces the per-call `Z / lengthscale` and `sum(square(Z / l), 1)` of
 * GPy Stationary._scaled_dist / _unscaled_dist (reached from edrgp/gp_model/base.py:69,222).
 * Builds, once per hyper-parameter set, the tiled device buffer every contraction kernel streams
 * through shared memory with one bulk copy per 32-point tile:
 *     tile t: 32 rows of  z_j / l^2  padded to a row stride of (dp + 2) doubles,
 *             then 32 x  -0.5 * ||z_j / l||^2,  then 32 x coefficient c_j
 * where dp = d rounded up to a multiple of 16 and padded points carry c_j = 0.
 * coef may be NULL (all ones: plain kernel entries) or a length-m device vector (sf2 * alpha_j for
 * the gradient kernel).
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_pack_bytes(int m, int d);
int edrgp_pack_inducing(const double* Z, const double* ell, const double* coef, double coef_scale,
                        int m, int d, double* pack, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  cross-covariance  Kfu = sf2 * exp(-0.5 * clip(|x/l|^2 + |z/l|^2 - 2 (x/l).(z/l), 0)).
 * Replaces GPy RBF.K(X, Z) (edrgp/gp_model/base.py:69 through VarDTC.inference, and :187).
 * Kfu is (n, ldk) row-major with ldk >= m.  Also optionally accumulates b += Kfu^T y (y may be
 * NULL).  b must be zeroed by the caller.
 * ------------------------------------------------------------------------------------------- */
int edrgp_kuf(const double* X, int64_t n, int d, const double* pack, int m, double sf2,
              double* Kfu, int64_t ldk, const double* y, double* b, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1+K4+K5 fused  posterior-mean gradients and their outer product.
 *   G_iq = scale * sum_j K_ij alpha_j (z_jq - x_iq) / l_q^2      (zero where clip(r^2) == 0)
 *   C    = G^T G  (d x d)
 * Replaces GPy GP.predictive_gradients -> Stationary.gradients_X(alpha^T, X, Z) reached from
 * edrgp/gp_model/base.py:222, and the Gram matrix behind SVDTransformer.fit
 * (edrgp/utils.py:140; eigh(G^T G) == right singular vectors of G).
 * `pack` must have been built with coef = alpha and coef_scale = sf2 * scale.
 * G may be NULL (gradients never leave the chip).  C may be NULL.  C is OVERWRITTEN with this
 * call's sum (reduction over CTAs is deterministic).  Requires d <= 128.
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_grad_gram_workspace_bytes(int d);
int edrgp_grad_gram(const double* X, int64_t n, int d, const double* pack, int m,
                    double* G, double* C, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDRGP_B200_H */
