// Internal launcher prototypes shared between the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edrgp {

cudaError_t launch_pack(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                        double* pack, cudaStream_t st);
bool grad_gram_fused(int d);
cudaError_t launch_grad_gram(const double* X, int64_t n, int d, const double* pack, int m, double* G, double* C,
                             double* Cpart, int sms, cudaStream_t st);
cudaError_t launch_kuf(const double* X, int64_t n, int d, const double* pack, int m, double sf2, double* Kfu,
                       int64_t ldk, const double* y, double* b, int sms, cudaStream_t st);

}  // namespace edrgp
