// extern "C" entry points of libedrgp_b200.so (declared in include/edrgp_b200.h).
#include <cstdio>
#include <cstdarg>
#include "../../include/edrgp_b200.h"
#include "common.cuh"
#include "launch.h"

namespace {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  return fail(EDRGP_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}
int sm_count_cached() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != dev_cached) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    dev_cached = dev;
  }
  return sms;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

extern "C" {

int edrgp_version(void) { return 100; }
const char* edrgp_last_error(void) { return g_err; }
int edrgp_sm_count(void) { return sm_count_cached(); }

size_t edrgp_pack_bytes(int m, int d) {
  if (m <= 0 || d <= 0) return 0;
  const int dp = edrgp::padded_dim(d);
  const size_t mtiles = (size_t)(m + edrgp::MT - 1) / edrgp::MT;
  return ((size_t)dp + mtiles * edrgp::pack_tile_doubles(dp)) * sizeof(double);
}

int edrgp_pack_inducing(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                        double* pack, void* stream) {
  if (!Z || !ell || !pack || m <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "pack_inducing: bad argument");
  if (!aligned16(pack)) return fail(EDRGP_ERR_ARG, "pack_inducing: pack must be 16-byte aligned");
  cudaError_t e = edrgp::launch_pack(Z, ell, coef, coef_scale, m, d, pack, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "pack_inducing");
}

static int check_x(const char* who, const double* X, int64_t n, int d, const double* pack, int m) {
  if (!X || !pack || n <= 0 || d <= 0 || m <= 0) return fail(EDRGP_ERR_ARG, "%s: bad argument", who);
  if (d > 128) return fail(EDRGP_ERR_UNSUPPORTED, "%s: d=%d > 128 needs the unfused path", who, d);
  if (d & 1) return fail(EDRGP_ERR_ARG, "%s: d must be even (pad X with a zero column)", who);
  if (!aligned16(X) || !aligned16(pack)) return fail(EDRGP_ERR_ARG, "%s: X and pack must be 16-byte aligned", who);
  return EDRGP_OK;
}

int edrgp_kuf(const double* X, int64_t n, int d, const double* pack, int m, double sf2, double* Kfu, int64_t ldk,
              const double* y, double* b, void* stream) {
  int rc = check_x("kuf", X, n, d, pack, m);
  if (rc) return rc;
  if (Kfu && (ldk < m || (ldk & 1) || !aligned16(Kfu))) return fail(EDRGP_ERR_ARG, "kuf: ldk must be even, >= m; Kfu 16-byte aligned");
  if ((y == nullptr) != (b == nullptr)) return fail(EDRGP_ERR_ARG, "kuf: y and b go together");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "kuf: no CUDA device");
  cudaError_t e = edrgp::launch_kuf(X, n, d, pack, m, sf2, Kfu, ldk, y, b, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "kuf");
}

size_t edrgp_grad_gram_workspace_bytes(int d) {
  const int dp = edrgp::padded_dim(d);
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  return (size_t)sms * dp * dp * sizeof(double);
}

int edrgp_grad_gram(const double* X, int64_t n, int d, const double* pack, int m, double* G, double* C,
                    void* workspace, void* stream) {
  int rc = check_x("grad_gram", X, n, d, pack, m);
  if (rc) return rc;
  if (G && !aligned16(G)) return fail(EDRGP_ERR_ARG, "grad_gram: G must be 16-byte aligned");
  if (C && !workspace) return fail(EDRGP_ERR_ARG, "grad_gram: C needs a workspace");
  if (C && !edrgp::grad_gram_fused(d))
    return fail(EDRGP_ERR_UNSUPPORTED, "grad_gram: fused Gram needs d <= 64; write G and call edrgp_syrk");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "grad_gram: no CUDA device");
  cudaError_t e = edrgp::launch_grad_gram(X, n, d, pack, m, G, C, (double*)workspace, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "grad_gram");
}

}  // extern "C"
