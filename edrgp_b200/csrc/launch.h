// Internal launcher prototypes shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edrgp {

// every kernel launch of the library is counted (edrgp_launch_count): bench.py reports it
void count_launch();

// dev_scale (optional): one double on the device multiplied into coef_scale inside the kernel
cudaError_t launch_pack(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                        double* pack, cudaStream_t st, const double* dev_scale = nullptr);
bool grad_gram_fused(int d);
cudaError_t launch_grad_gram(const double* X, int64_t n, int d, const double* pack, int m, double* G, double* C,
                             double* Cpart, int sms, cudaStream_t st);
cudaError_t launch_grad_gram_cached(const double* X, int64_t ldx, int64_t n, int d, const double* Kin, int64_t ldk,
                                    double sf2, const double* pack, int m, double* G, int64_t ldg, double* C,
                                    double* Cpart, int sms, cudaStream_t st);
cudaError_t launch_kuf(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int m, double sf2,
                       double* Kfu, int64_t ldk, int mul, const double* y, double* b, double* mu, int sms,
                       cudaStream_t st, int linear = 0, unsigned int* nonfinite_flag = nullptr);

// C (+)= A^T B (sym: B = A, upper tiles mirrored; optional y: bout[0..ka) (+)= A^T y, bout[ka] (+)= y^T y)
size_t gemm_tn_workspace_bytes(int64_t n, int ka, int kb, int sym, int sms);
cudaError_t launch_gemm_tn(const double* A, int64_t lda, int ka, const double* B, int64_t ldb, int kb, int64_t n,
                           int sym, const double* y, double* C, int64_t ldc, double* bout, int accumulate,
                           double* workspace, int sms, cudaStream_t st);
cudaError_t launch_potrf(double* A, int m, int64_t ld, int* info, cudaStream_t st);
cudaError_t launch_posv(double* A, int m, int64_t ld, double* L, int64_t ldl, double* rhs, double* x, int* info,
                        cudaStream_t st);
cudaError_t launch_trsm(const double* L, int m, int64_t ldl, double* B, int nrhs, int64_t ldb, int trans,
                        cudaStream_t st);
cudaError_t launch_kmm_fix(double* K, int m, int64_t ld, double sf2, double jitter, cudaStream_t st);
cudaError_t launch_solve(double* Kmm, const double* P, const double* b, int m, double beta, double* Bmat,
                         double* alpha, double* cvec, double* scalars, int* info, double* workspace,
                         cudaStream_t st);
cudaError_t launch_vfe_grad_small(const double* LB, const double* Lm, const double* Bmat, const double* c, int m,
                                  double beta, double* Msym, int64_t ldm, double* Dsym, double* sumAE,
                                  double* workspace, cudaStream_t st);
size_t eigh_workspace_doubles(int d);
cudaError_t launch_eigh(double* A, int d, double* V, double* evals, double* comps, int* sweeps, cudaStream_t st);

size_t col_moments_workspace_bytes(int d, int sms);
cudaError_t launch_col_moments(const double* X, int64_t n, int d, const double* shift, const double* weight,
                               double* out, int accumulate, double* workspace, int sms, cudaStream_t st);
size_t weights_workspace_bytes(int64_t n, int m);
cudaError_t launch_weights(const double* K, int64_t n, int m, int64_t ldk, const double* M, int64_t ldm, const double* y,
                           const double* alpha, double c_ya, double c_km, double* T, int64_t ldt, double* rowsum,
                           double* colsum, int accumulate, double* workspace, cudaStream_t st);
cudaError_t launch_count_nonfinite(const double* X, int64_t total, unsigned int* count, int sms, cudaStream_t st);
cudaError_t launch_standardize(const double* X, int64_t n, int d, const double* mean, const double* scale, double* out,
                               int sms, cudaStream_t st);
cudaError_t launch_project(const double* X, int64_t n, int d, const double* V, int k, double* out, int sms,
                           cudaStream_t st);

// TF32-split cross-covariance (tcgen05): tf32.cu
size_t pack_tf32_bytes(int m, int d);
cudaError_t launch_pack_tf32(const double* Z, const double* ell, int m, int d, void* pack, cudaStream_t st);
cudaError_t launch_kuf_tf32(const double* X, int64_t ldx, int64_t n, int d, const double* ell, const void* pack, int m,
                            double sf2, double* K, int64_t ldk, int sms, cudaStream_t st);

size_t pack_grad_tf32_bytes(int m, int d);
cudaError_t launch_pack_grad_tf32(const double* Z, const double* ell, const double* coef, double coef_scale, int m,
                                  int d, void* pack, cudaStream_t st);
cudaError_t launch_grad_tf32(const double* X, int64_t ldx, int64_t n, int d, const double* Kin, int64_t ldk, double sf2,
                             const double* ell, const void* pack, int m, double* G, int64_t ldg, int sms,
                             cudaStream_t st);

size_t pack_weights_tf32_bytes(int m);
cudaError_t launch_pack_weights_tf32(const double* M, int64_t ldm, double scale, int m, void* pack, cudaStream_t st);
cudaError_t launch_weights_tf32(const double* K, int64_t n, int m, int64_t ldk, const void* pack, const double* y,
                                const double* alpha, double c_ya, double* T, int64_t ldt, double* rowsum, int sms,
                                cudaStream_t st);

// exact-product INT8 symmetric reduction (tcgen05 kind::i8): i8syrk.cu
size_t inducing_stats_i8_workspace_bytes(int64_t n, int m, int sms);
cudaError_t launch_inducing_stats_i8(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2,
                                     double* P, int64_t ldp, double* b_yy, int accumulate, void* workspace, int sms,
                                     cudaStream_t st);

cudaError_t launch_i8_block(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2, int first,
                            void* workspace, int sms, cudaStream_t st);
cudaError_t launch_i8_finish(int m, double sf2, int with_y, double* P, int64_t ldp, double* b_yy, int accumulate,
                             void* workspace, int sms, cudaStream_t st);

cudaError_t launch_dmma_probe(double* scratch, int iters, int sms, double* flops, cudaStream_t st);

}  // namespace edrgp
