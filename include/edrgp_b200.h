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
/* number of CUDA kernels this library has launched in this process so far */
uint64_t edrgp_launch_count(void);

/* FP64 tensor-pipe (DMMA) peak probe: enqueues a register-only mma.sync f64 kernel; scratch is a
 * device buffer of >= 256 doubles; *flops (host) receives the FP64 flops the launch executes.  The
 * caller times it with CUDA events: the live roofline denominator of bench.py. */
int edrgp_fp64_probe(double* scratch, int iters, double* flops, void* stream);

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
/* dev_scale (may be NULL): one double ON THE DEVICE multiplied into coef_scale inside the kernel -- std(y) of
 * the target normaliser, which newer GPy multiplies into the Jacobian, without reading it back first. */
int edrgp_pack_inducing(const double* Z, const double* ell, const double* coef, double coef_scale,
                        const double* dev_scale, int m, int d, double* pack, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  cross-covariance  Kfu = sf2 * exp(-0.5 * clip(|x/l|^2 + |z/l|^2 - 2 (x/l).(z/l), 0)).
 * Replaces GPy RBF.K(X, Z) (edrgp/gp_model/base.py:69 through VarDTC.inference, and :187).
 * Kfu is (n, ldk) row-major with ldk >= m, ldk even; may be NULL.  Optionally accumulates
 * b += Kfu^T y (y, b may be NULL; b must be zeroed by the caller; atomics -- the deterministic
 * route is edrgp_inducing_stats) and writes mu_i = sum_j Kfu_ij coef_j (mu may be NULL): with
 * coef = alpha in the pack this is the posterior mean K(x, Z) alpha of GPy Posterior._raw_predict
 * (edrgp/gp_model/base.py:187).  The stored entries never carry the pack coefficient.
 * X is (n, ldx) with ldx even, >= d: only the first d columns from the pointer are used, so a block
 * of features of a wider matrix can be passed (d <= 128 per call).  multiply != 0 multiplies the
 * entries already in Kfu by this call's factor: exp(-r^2/2) factorises over feature blocks, which is
 * how d > 128 is evaluated (every block but the first with multiply = 1; the variance sf2, y, b
 * and mu only on the last block, the others with sf2 = 1).
 * nonfinite_flag (device, may be NULL) is set to 1 when a row holds a NaN or an Inf -- its scaled norm, which
 * the kernel forms anyway, is then not finite: the non-finite scan of sklearn's check_X_y
 * (edrgp/gp_model/base.py:87) at no extra pass over X.  The caller zeroes it.
 * ------------------------------------------------------------------------------------------- */
int edrgp_kuf(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int m, double sf2,
              double* Kfu, int64_t ldk, int multiply, const double* y, double* b, double* mu,
              unsigned int* nonfinite_flag, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1 in the TF32-split mode ("tf32x3"): the same cross-covariance as edrgp_kuf -- GPy RBF.K(X, Z),
 * edrgp/gp_model/base.py:69,187 -- with the (x/l).(z/l) contraction on the tcgen05 tensor cores.
 * Every FP64 operand is split into two TF32 values (22 significant bits) and three products are
 * accumulated in FP32 tensor memory; norms are reduced in FP64, the exponential is ex2.approx in
 * FP32, the entries are stored as FP64.  Entries agree with edrgp_kuf to ~1e-6 relative for
 * standardised inputs (the mode's contract is 1e-4).  d even, d <= 64; X (n, ldx) with ldx even.
 *   edrgp_pack_inducing_tf32: builds, once per hyper-parameter set, the device image the kernel
 *       streams: -|z/l|^2 log2(e) / 2 per inducing point (FP32), then per 128 points the hi and lo
 *       TF32 matrices of Z / l in the 128-byte-swizzled K-major layout tcgen05.mma reads.
 *   edrgp_kuf_tf32x3: Kfu (n, ldk) = sf2 exp(-r^2 / 2), r^2 clipped at 0; ldk >= m, even.
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_pack_tf32_bytes(int m, int d);
int edrgp_pack_inducing_tf32(const double* Z, const double* ell, int m, int d, void* pack, void* stream);
int edrgp_kuf_tf32x3(const double* X, int64_t ldx, int64_t n, int d, const double* ell, const void* pack,
                     int m, double sf2, double* Kfu, int64_t ldk, void* stream);

/* K4 in the TF32-split mode: posterior-mean gradients from the STORED cross-covariance block,
 *     G_iq = sum_j K_ij c_j (z_jq - x_iq) / l_q^2,   c = alpha * scale
 * (GPy GP.predictive_gradients -> Stationary.gradients_X, edrgp/gp_model/base.py:222) as ONE contraction
 * K (n, m) x B (m, 1 + d) on tcgen05 (B_j0 = c_j: the row sums; B_j,1+q = c_j z_jq / l_q^2), K split into
 * hi / lo TF32 on the fly, FP32 accumulation in tensor memory, the rowsum * x / l^2 term subtracted in
 * FP64.  GPy's dropped pairs (r = 0, entries equal to sf2) are not special-cased: their terms cancel
 * between the contraction and the correction to the mode's rounding level (sf2 is accepted for symmetry
 * with edrgp_grad_gram_cached).  d even, d <= 64; ldk, ldx, ldg even.  The Gram matrix G^T G then runs on the FP64 reduction (edrgp_syrk).
 *   edrgp_pack_grad_tf32: builds B's device image (per 32 inducing points: hi | lo, 128-byte swizzled). */
size_t edrgp_pack_grad_tf32_bytes(int m, int d);
int edrgp_pack_grad_tf32(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                         void* pack, void* stream);
int edrgp_grad_tf32x3(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk, double sf2,
                      const double* ell, const void* pack, int m, double* G, int64_t ldg, void* stream);

/* The weights of the hyper-parameter gradient in the TF32-split mode (edrgp_weights is the FP64 form):
 *     T = K o (c_ya y alpha^T + K M'),  rowsum_i = sum_j T_ij,        M' = scale * M (symmetric), m <= 512
 * -- dL/dKfu applied to the stored cross-covariance, GPy VarDTC dL_dpsi1 / dL_dpsi2 reached from
 * model.optimize (edrgp/gp_model/base.py:69) -- with the n x m x m contraction K M' on tcgen05 (K split
 * into hi / lo TF32 on the fly, M' prepacked by edrgp_pack_weights_tf32, FP32 accumulation in tensor
 * memory) and the elementwise part in FP64 (the contraction kernel writes c_ya y alpha^T + K M' into T, a
 * streaming pass multiplies by K and reduces the rows).  y, alpha may be NULL (c_ya term dropped); rowsum (n)
 * may be NULL; T (n, ldt) is required; ldk, ldt even.  Column sums of T need no pass over it:
 * sum_i T_ij = c_ya alpha_j (K^T y)_j + sum_k (K^T K)_jk M'_kj. */
size_t edrgp_pack_weights_tf32_bytes(int m);
int edrgp_pack_weights_tf32(const double* M, int64_t ldm, double scale, int m, void* pack, void* stream);
int edrgp_weights_tf32x3(const double* K, int64_t n, int m, int64_t ldk, const void* pack, const double* y,
                         const double* alpha, double c_ya, double* T, int64_t ldt, double* rowsum, void* stream);

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

/* The same gradients and Gram matrix from a STORED cross-covariance block (the one edrgp_kuf wrote
 * for the statistics pass: Kfu (n, ldk), entries sf2 exp(-r^2/2)) instead of recomputing it: only
 * the W Z contraction, the row sums and G^T G remain.  `pack` must have been built with coef = alpha
 * and coef_scale = scale (the entries already carry sf2); entries exactly equal to sf2 are the
 * pairs with clipped r^2 == 0 that GPy's _inv_dist drops.  Requires d <= 64 per call: X (n, ldx) and
 * G (n, ldg) may be feature blocks of wider matrices (the gradient of feature q only needs the full
 * Kfu and column q of X and Z), which is how d > 64 is covered block by block; C is then the
 * block's own d x d Gram matrix (use edrgp_syrk on the assembled G for the full one).  Same workspace. */
int edrgp_grad_gram_cached(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk,
                            double sf2, const double* pack, int m, double* G, int64_t ldg, double* C,
                            void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2 / K5  tall-skinny reductions on the FP64 tensor pipe.  All matrices row-major, leading
 * dimensions even, base pointers 16-byte aligned; results are deterministic (fixed-order split-K
 * reduction, no atomics); accumulate != 0 adds to the output instead of overwriting it (row chunks).
 *
 *  edrgp_syrk            C (k, ldc) (+)= A^T A, full symmetric.  Replaces GPy tdot (dsyrk) and the
 *                        Gram matrix behind np.linalg.svd(G) when d > 64 (edrgp/utils.py:140).
 *  edrgp_inducing_stats  P (m, ldp) (+)= Kfu^T Kfu,  b_yy[0..m) (+)= Kfu^T y,  b_yy[m] (+)= y^T y:
 *                        the n-reduced statistics of GPy VarDTC.inference (psi2 = tdot(psi1^T),
 *                        psi1^T Y, trYYT) reached from edrgp/gp_model/base.py:69.
 *  edrgp_gemm_tn         C (ka, ldc) (+)= A^T B for A (n, lda), B (n, ldb): T^T X of the
 *                        hyper-parameter gradients (GPy Stationary.gradients_X w.r.t. Z and
 *                        update_gradients_full, reached from model.optimize, base.py:69).
 * workspace: edrgp_syrk_workspace_bytes(n, k) for the first two, edrgp_gemm_tn_workspace_bytes.
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_syrk_workspace_bytes(int64_t n, int k);
int edrgp_syrk(const double* A, int64_t n, int k, int64_t lda, double* C, int64_t ldc, int accumulate,
               void* workspace, void* stream);
int edrgp_inducing_stats(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double* P,
                         int64_t ldp, double* b_yy, int accumulate, void* workspace, void* stream);
size_t edrgp_gemm_tn_workspace_bytes(int64_t n, int ka, int kb);
int edrgp_gemm_tn(const double* A, int64_t lda, int ka, const double* B, int64_t ldb, int kb, int64_t n,
                  double* C, int64_t ldc, int accumulate, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * T = Kfu o (c_ya * y alpha^T + c_km * Kfu M) for a stored cross-covariance block Kfu (n, ldk) and an
 * (m, ldm) matrix M; optional outputs T (n, ldt), rowsum (n) = T 1, colsum (m) (+)= T^T 1.
 *  - M = dL/dpsi2, c_ya = beta, c_km = 2:  T = Kfu o dL/dKfu of GPy VarDTC.inference, the weight
 *    matrix of Stationary.update_gradients_full / gradients_X (model.optimize,
 *    edrgp/gp_model/base.py:69);
 *  - M = woodbury_inv, c_ya = 0 (y = alpha = NULL), c_km = 1: rowsum_i = k_i^T W k_i of the
 *    predictive variance (GPy Posterior._raw_predict, edrgp/gp_model/base.py:206).
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_weights_workspace_bytes(int64_t n, int m);
int edrgp_weights(const double* Kfu, int64_t n, int m, int64_t ldk, const double* M, int64_t ldm,
                  const double* y, const double* alpha, double c_ya, double c_km, double* T, int64_t ldt,
                  double* rowsum, double* colsum, int accumulate, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kuu = K(Z, Z) with the diagonal forced to sf2 + jitter (GPy Stationary._unscaled_dist zeroes the
 * diagonal distance; VarDTC adds const_jitter = 1e-8).  Zp is the (m, d) inducing matrix with d
 * even (as stored by the host pack), `pack` built with coef = NULL.  Kmm is (m, ldk) row-major
 * with ldk even and >= m (rows are written with 16-byte stores).  Zp is (m, ldz); like edrgp_kuf it
 * can be called per feature block (multiply), with finish != 0 on the last block only (diagonal,
 * jitter and symmetrisation).
 * ------------------------------------------------------------------------------------------- */
int edrgp_kmm(const double* Zp, int64_t ldz, const double* pack, int m, int d, double sf2, double jitter,
              double* Kmm, int64_t ldk, int multiply, int finish, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  the m x m solve chain of GPy VarDTC.inference, driven by the n-reduced statistics
 * (edrgp/gp_model/base.py:69):
 *     Lm = chol(Kmm);  A = beta Lm^-1 P Lm^-T;  LB = chol(I + A);
 *     c = LB^-1 Lm^-1 (beta b);  alpha = Lm^-T LB^-T c            (GPy woodbury_vector)
 * In:  Kmm (m, m) -- OVERWRITTEN by Lm (lower triangle);  P (m, m);  b (m);  beta = 1 / noise.
 * Out: LB (m, m) lower triangle;  alpha (m);  c (m);
 *      scalars[0] = tr(A), scalars[1] = sum log diag(LB), scalars[2] = c^T c;
 *      info[0], info[1]: 0 or 1 + index of the first non-positive pivot of chol(Kmm) / chol(I + A).
 * workspace: edrgp_solve_workspace_bytes(m).
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_solve_workspace_bytes(int m);
int edrgp_solve(double* Kmm, const double* P, const double* b, int m, double beta, double* LB,
                double* alpha, double* c, double* scalars, int* info, void* workspace, void* stream);

/* Lower Cholesky factor in place (GPy jitchol -> LAPACK dpotrf): A (m, ld) row-major, the strict
 * upper triangle is left untouched; info[0] = 0 or 1 + index of the first non-positive pivot.
 * With A = Kuu + beta P its factor LS = Lm LB gives the posterior weights directly,
 * alpha = LS^-T LS^-1 (beta b): the fixed-hyper-parameter sweep needs nothing else of the chain. */
int edrgp_potrf(double* A, int m, int64_t ld, int* info, void* stream);

/* The m x m part of the hyper-parameter gradients of the bound (GPy VarDTC.inference: _compute_dL_dpsi and
 * dL_dKmm, reached from edrgp/gp_model/base.py:69 on every optimiser evaluation), from the factors edrgp_solve left:
 *     E = LB^-T (I + c c^T) LB^-1;   dL_dpsi2 = beta/2 Lm^-T (I - E) Lm^-1;   dL_dKmm = Lm^-T (I - E/2 - B/2) Lm^-1
 * In:  LB, Lm (m, m) lower factors;  B (m, m) = I + A as kept in the first m^2 doubles of edrgp_solve's workspace;
 *      c (m).   Out: Msym (m, ldm) = (dL_dpsi2 + dL_dpsi2^T) / 2;  Dsym (m, m) = (dL_dKmm + dL_dKmm^T) / 2;
 *      sumAE[0] = sum (B - I) o E.   workspace: edrgp_vfe_grad_small_workspace_bytes(m). */
size_t edrgp_vfe_grad_small_workspace_bytes(int m);
int edrgp_vfe_grad_small(const double* LB, const double* Lm, const double* B, const double* c, int m, double beta,
                         double* Msym, int64_t ldm, double* Dsym, double* sumAE, void* workspace, void* stream);

/* A x = rhs by Cholesky with one right-hand side (LAPACK dposv, nrhs = 1): the posterior weights
 * alpha = (Kuu + beta P)^-1 beta b of the fixed-hyper-parameter sweep (GPy Posterior.woodbury_vector,
 * edrgp/gp_model/base.py:69,222) in one call.  One launch per 32-column blocked step (the diagonal
 * factor, the panel solves and the trailing update of a tile fused in one CTA), the right-hand side
 * carried as one more row of the matrix so that the forward substitution costs nothing, then one
 * backward substitution.  A (m, ld): lower triangle DESTROYED;  L (m, ldl), a different buffer:
 * receives the factor (lower triangle, the rest untouched);  rhs (m): destroyed;  x (m), a different
 * buffer: the solution.  rhs = x = NULL factors only.  info[0] as in edrgp_potrf. */
int edrgp_posv(double* A, int m, int64_t ld, double* L, int64_t ldl, double* rhs, double* x, int* info,
               void* stream);

/* lower-triangular solves with a factor from edrgp_solve: trans = 0: L X = B, 1: L^T X = B;
 * B (m, nrhs) row-major, overwritten.  (GPy dtrtrs.) */
int edrgp_trsm(const double* L, int m, double* B, int nrhs, int trans, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  symmetric eigendecomposition of the d x d positive semi-definite EDR matrix C = G^T G
 * (one-sided cyclic Jacobi on W = C V; eigenvalues as Rayleigh quotients).
 * Replaces np.linalg.svd(G) in SVDTransformer.fit (edrgp/utils.py:140): comps rows are the right
 * singular vectors of G, evals = S^2, descending.  C is left intact.
 * workspace: edrgp_eigh_workspace_bytes(d).  sweeps (TWO device ints, may be NULL; the caller zeroes them):
 * [0] receives the number of Jacobi sweeps used (60 = the limit: not converged), [1] a status word that is
 * non-zero when the multi-SM solver's grid barrier timed out (results are then unusable).  d <= 116 runs as one CTA in shared memory (d <= 64: the solver records its rotations and a
 * second kernel replays them on V across SMs); larger d runs all sweeps inside one persistent multi-SM kernel
 * with a grid barrier per round-robin step.  Every size only enqueues.
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_eigh_workspace_bytes(int d);
int edrgp_eigh(double* C, int d, double* evals, double* comps, int* sweeps, void* workspace,
               void* stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming row kernels around the path (HBM-bound).
 *  edrgp_col_moments: out[q] (+)= sum_i w_i (x_iq - shift_q), out[d + q] (+)= sum_i w_i (x_iq - shift_q)^2
 *      (shift, weight may be NULL; weight = rowsum(T) gives the lengthscale-gradient moment).  Two calls give the mean and the centred second moment of
 *      StandardScaler (edrgp/edr.py:161-162) and of GPy's Standardize on y; the sums are what an
 *      n-sharded run all-reduces.  d <= 512.
 *  edrgp_standardize: out = (X - mean) / scale, elementwise by column (in place allowed).
 *  edrgp_project:     out (n, k) = X (n, d) V^T, V (k, d) row-major: EDR.transform
 *      (edrgp/edr.py:261-289, edrgp/base.py:462).
 * ------------------------------------------------------------------------------------------- */
 /* edrgp_count_nonfinite: count[0] += #(NaN or Inf entries) of a dense buffer of `total` doubles:
 *      the non-finite scan of sklearn's check_X_y / check_array (edrgp/gp_model/base.py:87,105). */
int edrgp_count_nonfinite(const double* X, int64_t total, unsigned int* count, void* stream);
 /* edrgp_project_dmma: the same projection on the FP64 tensor pipe for many components (k > 8):
 *      out (n, ldo) = X (n, ldx)[:, :d] V^T with V given as an inducing pack built from V with unit
 *      lengthscales (edrgp_pack_inducing(V, ones, NULL, 1, k, d, ...)); d <= 128, ldo even. */
int edrgp_project_dmma(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int k,
                       double* out, int64_t ldo, void* stream);
size_t edrgp_col_moments_workspace_bytes(int d);
int edrgp_col_moments(const double* X, int64_t n, int d, const double* shift, const double* weight,
                      double* out, int accumulate, void* workspace, void* stream);
int edrgp_standardize(const double* X, int64_t n, int d, const double* mean, const double* scale,
                      double* out, void* stream);
int edrgp_project(const double* X, int64_t n, int d, const double* V, int k, double* out,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2 on the INT8 tensor cores (tcgen05.mma.kind::i8): the same statistics as edrgp_inducing_stats --
 * P (+)= Kfu^T Kfu, b_yy[:m] (+)= Kfu^T y, b_yy[m] (+)= y^T y (GPy tdot(psi1) / psi1^T Y in VarDTC.inference,
 * edrgp/gp_model/base.py:69) -- from EXACT integer products: Kfu entries, which must lie in [0, sf2], are rounded to
 * 48 fraction bits of K / (4 sf2) and written as six signed radix-256 digits (balanced, [-128, 127]); digit products
 * accumulate in 32-bit integers in tensor memory, pairs of equal weight share an accumulator, and
 * P = 16 sf2^2 sum_g 2^-8(g+2) A_g.  Both error terms (rounding at 2^-49, dropped pairs of weight 2^-64) are
 * zero-mean, so P agrees with the FP64 reduction to FP64 rounding level (~1e-15 of max |P|).
 * Kfu (n, ldk) row-major as written by edrgp_kuf with THIS sf2; y, b_yy may both be NULL; m <= 2048.
 * workspace: edrgp_inducing_stats_i8_workspace_bytes(n, m) (the digit planes of the block: 6 n m bytes), 128-byte
 * aligned (the planes are written in 32-byte sectors and read with 16 KB bulk copies).
 * ------------------------------------------------------------------------------------------- */
size_t edrgp_inducing_stats_i8_workspace_bytes(int64_t n, int m);
int edrgp_inducing_stats_i8(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2, double* P,
                            int64_t ldp, double* b_yy, int accumulate, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The fixed-hyper-parameter EDR sweep as composite calls: everything between two collectives is
 * enqueued by ONE call (a rank whose row shard takes a few milliseconds must not wait for its host
 * between kernels).  Together they replace one estimator.fit + predict_gradient + SVDTransformer.fit
 * of the reference loop (edrgp/base.py:435-466; edrgp/gp_model/base.py:46-70,208-222;
 * edrgp/utils.py:123-157) at fixed hyper-parameters, for even d <= 64 with the whole Kfu (n, ldk)
 * kept in HBM.  All calls share one workspace of edrgp_fixed_layout(...) bytes whose regions
 * (offsets in doubles, written to offsets[EDRGP_FS_NREGIONS]) the caller may read -- and all-reduce in
 * place, which is the only thing a multi-rank caller does between the calls:
 *
 *   edrgp_fixed_begin      pack(Z, l) -> Kfu block 0 (no targets needed: the device is busy at once)
 *                          -> this rank's [n_r, pivot, S1, S2] of the targets into row `rank` of TABLE
 *       [sum TABLE over ranks: every rank fills only its own row]
 *   edrgp_fixed_stats      mean / std(y) over all ranks (GPy Standardize; normalize = 0: y is used as it is and only
 *                          N is refreshed), TAIL = [flags | N | mean | std], then P, b, y^T y over all row blocks
 *                          into STATS
 *       [sum STATS over ranks]
 *   edrgp_fixed_posterior  S = Kuu + jitter I + beta P;  alpha = S^-1 beta b  (Cholesky; info -> TAIL)
 *   edrgp_fixed_grad       pack(Z, l, alpha * coef_scale [* dev_scale[0], e.g. std(y) = TAIL[3]]) -> gradients from the stored Kfu
 *                          (G (n, ldg) optional) -> C = G^T G into RESULT, TAIL copied behind it
 *       [sum C over ranks]
 *   edrgp_fixed_eigh       eigh(C): RESULT = evals (d) | components (d x d) | C (d x d) | TAIL copy (4)
 *
 * TAIL word 0 holds two 32-bit integers: the non-finite flag of edrgp_kuf and the info of edrgp_posv.
 * h2d (may be NULL): an edrgp_h2d_open handle whose rows / side payload are X / y still on their way from the host;
 * begin and stats then make the stream wait for each row block (chunk_rows rows) right before the kernels that read it.
 * ------------------------------------------------------------------------------------------- */
enum {
  EDRGP_FS_PACK_K = 0, EDRGP_FS_PACK_G, EDRGP_FS_YT, EDRGP_FS_STATS, EDRGP_FS_TABLE, EDRGP_FS_S, EDRGP_FS_L,
  EDRGP_FS_RHS, EDRGP_FS_ALPHA, EDRGP_FS_SCRATCH, EDRGP_FS_TAIL, EDRGP_FS_RESULT, EDRGP_FS_NREGIONS
};
size_t edrgp_fixed_layout(int64_t n, int d, int m, int64_t chunk_rows, int world, int64_t* offsets);

/* Statistics route of edrgp_fixed_stats: 0 = the FP64 DMMA reduction (edrgp_inducing_stats; default), 1 = the
 * exact-product INT8 route (edrgp_inducing_stats_i8; m <= 2048, otherwise the FP64 route runs).  Process-wide and
 * part of the workspace layout: set it BEFORE edrgp_fixed_layout sizes a workspace (route 1 adds room for one row
 * block's slices, 6 chunk_rows m bytes).  Initial value from the environment: EDRGP_STATS = fp64 | int8x6. */
int edrgp_set_stats_mode(int mode);
int edrgp_get_stats_mode(void);
int edrgp_fixed_begin(const double* X, int64_t ldx, int64_t n, int d, const double* y, const double* Z, int64_t ldz,
                      const double* ell, int m, double sf2, int64_t chunk_rows, double* Kfu, int64_t ldk, int rank,
                      int world, void* h2d, int64_t h2d_ahead, void* workspace, void* stream);
int edrgp_fixed_stats(const double* X, int64_t ldx, int64_t n, int d, const double* y, int m, double sf2,
                      int64_t chunk_rows, double* Kfu, int64_t ldk, int normalize, int world, void* h2d, int64_t h2d_ahead,
                      void* workspace, void* stream);
int edrgp_fixed_posterior(const double* Z, int64_t ldz, int64_t n, int d, int m, double sf2, double jitter, double beta,
                          int64_t chunk_rows, int world, void* workspace, void* stream);
int edrgp_fixed_grad(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk, const double* Z,
                     int64_t ldz, const double* ell, int m, double sf2, double coef_scale, const double* dev_scale,
                     double* G, int64_t ldg, int64_t chunk_rows, int world, void* workspace, void* stream);
int edrgp_fixed_eigh(int64_t n, int d, int m, int64_t chunk_rows, int world, void* workspace, void* stream);

/* -------------------------------------------------------------------------------------------
 * NVLink peer exchange for the sweep's three reductions (csrc/peer.cu).  No counterpart in the reference, which is a
 * single process (edrgp/base.py:435-466 runs one estimator over all rows): this is the n-sharding of the path, and it
 * replaces the three all-reduce calls a multi-rank caller otherwise issues between the composite calls above.
 *
 * Every rank allocates one exchange buffer (edrgp_peer_alloc: cudaMalloc + a 64-byte cudaIpcMemHandle the caller ships
 * to the other ranks by whatever channel it has), maps the others' (edrgp_peer_open) and binds the set to a sweep
 * workspace (edrgp_fixed_bind_peers: bases[r] = rank r's buffer as mapped HERE, bases[rank] = its own; bases = NULL
 * unbinds).  With a bound workspace
 *   edrgp_fixed_begin      also pushes this rank's moments row into every rank's table and raises its flag,
 *   edrgp_fixed_stats      waits for all rows first, writes the rank's partial {P, b, y^T y} into its exchange buffer and
 *                          raises its flag when the last block is done,
 *   edrgp_fixed_posterior  sums the partials of all ranks, in rank order, WHILE it assembles S = Kuu + beta P and beta b
 *                          (reads over NVLink; the sums are also left in the EDRGP_FS_STATS region),
 *   edrgp_fixed_reduce_gram (bound workspaces only) replaces C in the result block by its sum over the ranks,
 * and the caller issues NO collective of its own.  All ranks must make the same sequence of calls.  Every wait is
 * bounded (~10 s): a rank that never arrives sets bit 8 (0x100) of the flag word in EDRGP_FS_TAIL instead of hanging.
 * edrgp_peer_layout returns the buffer size in bytes for (m, d, world <= 16); offsets[4] (doubles): flags, table,
 * statistics, Gram matrix -- two copies of each payload, used alternately.
 * ------------------------------------------------------------------------------------------- */
#define EDRGP_PEER_HANDLE_BYTES 64
size_t edrgp_peer_layout(int m, int d, int world, int64_t* offsets);
int edrgp_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle);
int edrgp_peer_open(const void* ipc_handle, void** dev_ptr);
int edrgp_peer_close(void* dev_ptr);
int edrgp_peer_free(void* dev_ptr);
int edrgp_fixed_bind_peers(void* workspace, void* const* bases, int rank, int world, int m, int d);
int edrgp_fixed_reduce_gram(int64_t n, int d, int m, int64_t chunk_rows, int world, void* workspace, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Host rows -> device in blocks, overlapped with the kernels that consume them.  Replaces the implicit
 * "the arrays are already where the arithmetic runs" of the reference's fit(X, y) (edrgp/gp_model/base.py:46-91:
 * check_X_y hands GPy a C-contiguous float64 host ndarray).
 *   edrgp_h2d_open   starts the transfer of `rows` rows of `row_bytes` from `host` (contiguous) to `dev` (row pitch
 *                    dst_pitch >= row_bytes) in blocks of block_rows rows, ordered after the work already enqueued on
 *                    order_after_stream.  A pinned / registered source is copied with one cudaMemcpyAsync per block;
 *                    an ordinary (pageable) source is staged by `threads` host threads through a ring of `slots`
 *                    pinned blocks (kept for the next transfer), each block's DMA enqueued by the thread that
 *                    completes it.  side_host / side_dev / side_bytes (may be NULL / 0): one more contiguous buffer
 *                    -- the targets -- that travels on the same copy stream right behind the first block
 *                    (edrgp_h2d_wait_side makes a stream wait for it).  Returns a handle, or NULL (edrgp_last_error).
 *   edrgp_h2d_wait   blocks the calling host thread until the blocks covering rows [0, upto_row) have been ENQUEUED,
 *                    then makes consumer_stream wait for them (no device synchronisation).  A pinned source is
 *                    enqueued here, up to ahead_rows beyond upto_row (the copy engine serves its queue in order: a
 *                    transfer enqueued all at once would delay every small upload the caller makes next).
 *   edrgp_h2d_staged 1 when the source was pageable and goes through the ring.
 *   edrgp_h2d_close  joins the threads, waits for the copies and releases the handle; `host` must stay valid until then.
 * ------------------------------------------------------------------------------------------- */
void* edrgp_h2d_open(const void* host, void* dev, int64_t rows, size_t row_bytes, size_t dst_pitch, int64_t block_rows,
                     int threads, int slots, void* order_after_stream, const void* side_host, void* side_dev,
                     size_t side_bytes);
int edrgp_h2d_wait_side(void* handle, void* consumer_stream);
int edrgp_h2d_wait(void* handle, int64_t upto_row, int64_t ahead_rows, void* consumer_stream);
int edrgp_h2d_staged(void* handle);
int edrgp_h2d_close(void* handle);

/* Optional per-stage timing of the composite calls (measurement aid; nothing on the product path needs it):
 * between edrgp_timing_begin and edrgp_timing_end every stage a composite call launches is bracketed by a pair of
 * CUDA events on the caller's stream; _end waits for them and adds up milliseconds and launch groups per stage
 * (ms[EDRGP_STAGE_COUNT], count[EDRGP_STAGE_COUNT]).  EDRGP_STAGE_STATS spans exactly one launch of the symmetric
 * reduction P = Kfu^T Kfu per row block, which is what bench.py's roofline divides by. */
enum { EDRGP_STAGE_KUF = 0, EDRGP_STAGE_TARGETS, EDRGP_STAGE_STATS, EDRGP_STAGE_SOLVE, EDRGP_STAGE_GRAD, EDRGP_STAGE_EIGH,
       EDRGP_STAGE_COUNT };
int edrgp_timing_begin(void);
int edrgp_timing_end(double* ms, int* count);

#ifdef __cplusplus
}
#endif
#endif /* EDRGP_B200_H */
