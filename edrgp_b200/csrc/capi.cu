// extern "C" entry points of libedrgp_b200.so (declared in include/edrgp_b200.h).
#include <atomic>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "../../include/edrgp_b200.h"
#include "common.cuh"
#include "launch.h"
#include "sweep.h"
#include "h2d.h"
#include "peer.h"
#include <unordered_map>
#include <iterator>

namespace edrgp {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace edrgp

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

// ---- optional per-stage CUDA-event timing of the composite calls (edrgp_timing_begin / _end): the benchmark's
// roofline needs the duration of individual kernels that one composite call launches back to back
struct StageTimer {
  bool on = false;
  std::vector<cudaEvent_t> pool;            // reused across sessions
  size_t used = 0;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return nullptr; pool.push_back(e); }
    return pool[used++];
  }
};
StageTimer g_timer;
std::mutex g_timer_mu;
struct StageScope {                         // records an event pair around the launches of one stage
  cudaStream_t st; int idx = -1;
  StageScope(int stage, cudaStream_t s) : st(s) {
    if (!g_timer.on) return;
    std::lock_guard<std::mutex> lk(g_timer_mu);
    cudaEvent_t a = g_timer.get(), b = g_timer.get();
    if (!a || !b) return;
    cudaEventRecord(a, st);
    idx = (int)g_timer.spans.size();
    g_timer.spans.push_back({stage, a, b});
  }
  ~StageScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_timer_mu);
    cudaEventRecord(g_timer.spans[idx].b, st);
  }
};
}  // namespace

extern "C" {

int edrgp_version(void) { return 100; }
const char* edrgp_last_error(void) { return g_err; }
int edrgp_sm_count(void) { return sm_count_cached(); }
uint64_t edrgp_launch_count(void) { return edrgp::g_launches.load(std::memory_order_relaxed); }

int edrgp_fp64_probe(double* scratch, int iters, double* flops, void* stream) {
  if (!scratch || iters <= 0) return fail(EDRGP_ERR_ARG, "fp64_probe: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "fp64_probe: no CUDA device");
  cudaError_t e = edrgp::launch_dmma_probe(scratch, iters, sms, flops, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "fp64_probe");
}

size_t edrgp_pack_bytes(int m, int d) {
  if (m <= 0 || d <= 0) return 0;
  const int dp = edrgp::padded_dim(d);
  const size_t mtiles = (size_t)(m + edrgp::MT - 1) / edrgp::MT;
  return ((size_t)dp + mtiles * edrgp::pack_tile_doubles(dp)) * sizeof(double);
}

int edrgp_pack_inducing(const double* Z, const double* ell, const double* coef, double coef_scale,
                        const double* dev_scale, int m, int d, double* pack, void* stream) {
  if (!Z || !ell || !pack || m <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "pack_inducing: bad argument");
  if (!aligned16(pack)) return fail(EDRGP_ERR_ARG, "pack_inducing: pack must be 16-byte aligned");
  cudaError_t e = edrgp::launch_pack(Z, ell, coef, coef_scale, m, d, pack, (cudaStream_t)stream, dev_scale);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "pack_inducing");
}

static int check_x(const char* who, const double* X, int64_t n, int d, const double* pack, int m) {
  if (!X || !pack || n <= 0 || d <= 0 || m <= 0) return fail(EDRGP_ERR_ARG, "%s: bad argument", who);
  if (d > 128) return fail(EDRGP_ERR_UNSUPPORTED, "%s: d=%d > 128 needs the unfused path", who, d);
  if (d & 1) return fail(EDRGP_ERR_ARG, "%s: d must be even (pad X with a zero column)", who);
  if (!aligned16(X) || !aligned16(pack)) return fail(EDRGP_ERR_ARG, "%s: X and pack must be 16-byte aligned", who);
  return EDRGP_OK;
}

int edrgp_kuf(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int m, double sf2, double* Kfu,
              int64_t ldk, int multiply, const double* y, double* b, double* mu, unsigned int* nonfinite_flag,
              void* stream) {
  int rc = check_x("kuf", X, n, d, pack, m);
  if (rc) return rc;
  if (ldx < d || (ldx & 1)) return fail(EDRGP_ERR_ARG, "kuf: ldx must be even and >= d");
  if (multiply && !Kfu) return fail(EDRGP_ERR_ARG, "kuf: multiply needs the Kfu buffer");
  if (Kfu && (ldk < m || (ldk & 1) || !aligned16(Kfu))) return fail(EDRGP_ERR_ARG, "kuf: ldk must be even, >= m; Kfu 16-byte aligned");
  if ((y == nullptr) != (b == nullptr)) return fail(EDRGP_ERR_ARG, "kuf: y and b go together");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "kuf: no CUDA device");
  if (!(sf2 > 0.0) || !(sf2 < 1e300)) return fail(EDRGP_ERR_ARG, "kuf: the kernel variance must be positive and finite");
  cudaError_t e = edrgp::launch_kuf(X, ldx, n, d, pack, m, sf2, Kfu, ldk, multiply, y, b, mu, sms, (cudaStream_t)stream,
                                    0, nonfinite_flag);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "kuf");
}

size_t edrgp_pack_tf32_bytes(int m, int d) {
  if (m <= 0 || d <= 0 || d > 64) return 0;
  return edrgp::pack_tf32_bytes(m, d);
}

int edrgp_pack_inducing_tf32(const double* Z, const double* ell, int m, int d, void* pack, void* stream) {
  if (!Z || !ell || !pack || m <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "pack_inducing_tf32: bad argument");
  if (d > 64) return fail(EDRGP_ERR_UNSUPPORTED, "pack_inducing_tf32: d=%d > 64 is outside the TF32-split mode", d);
  if (!aligned16(pack)) return fail(EDRGP_ERR_ARG, "pack_inducing_tf32: pack must be 16-byte aligned");
  cudaError_t e = edrgp::launch_pack_tf32(Z, ell, m, d, pack, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "pack_inducing_tf32");
}

int edrgp_kuf_tf32x3(const double* X, int64_t ldx, int64_t n, int d, const double* ell, const void* pack, int m,
                     double sf2, double* Kfu, int64_t ldk, void* stream) {
  if (!X || !ell || !pack || !Kfu || n <= 0 || d <= 0 || m <= 0) return fail(EDRGP_ERR_ARG, "kuf_tf32x3: bad argument");
  if (d > 64) return fail(EDRGP_ERR_UNSUPPORTED, "kuf_tf32x3: d=%d > 64 is outside the TF32-split mode", d);
  if (m > 4096) return fail(EDRGP_ERR_UNSUPPORTED, "kuf_tf32x3: m=%d > 4096", m);
  if ((d & 1) || ldx < d || (ldx & 1)) return fail(EDRGP_ERR_ARG, "kuf_tf32x3: d and ldx must be even, ldx >= d");
  if (ldk < m || (ldk & 1)) return fail(EDRGP_ERR_ARG, "kuf_tf32x3: ldk must be even and >= m");
  if (!(sf2 > 0.0)) return fail(EDRGP_ERR_ARG, "kuf_tf32x3: the kernel variance must be positive");
  if (!aligned16(X) || !aligned16(pack) || !aligned16(Kfu)) return fail(EDRGP_ERR_ARG, "kuf_tf32x3: X, pack and Kfu must be 16-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "kuf_tf32x3: no CUDA device");
  cudaError_t e = edrgp::launch_kuf_tf32(X, ldx, n, d, ell, pack, m, sf2, Kfu, ldk, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "kuf_tf32x3");
}

size_t edrgp_pack_grad_tf32_bytes(int m, int d) {
  if (m <= 0 || d <= 0 || d > 64) return 0;
  return edrgp::pack_grad_tf32_bytes(m, d);
}

int edrgp_pack_grad_tf32(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                         void* pack, void* stream) {
  if (!Z || !ell || !coef || !pack || m <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "pack_grad_tf32: bad argument");
  if (d > 64) return fail(EDRGP_ERR_UNSUPPORTED, "pack_grad_tf32: d=%d > 64 is outside the TF32-split mode", d);
  if (!aligned16(pack)) return fail(EDRGP_ERR_ARG, "pack_grad_tf32: pack must be 16-byte aligned");
  cudaError_t e = edrgp::launch_pack_grad_tf32(Z, ell, coef, coef_scale, m, d, pack, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "pack_grad_tf32");
}

int edrgp_grad_tf32x3(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk, double sf2,
                      const double* ell, const void* pack, int m, double* G, int64_t ldg, void* stream) {
  if (!X || !Kfu || !ell || !pack || !G || n <= 0 || d <= 0 || m <= 0) return fail(EDRGP_ERR_ARG, "grad_tf32x3: bad argument");
  if (d > 64) return fail(EDRGP_ERR_UNSUPPORTED, "grad_tf32x3: d=%d > 64 is outside the TF32-split mode", d);
  if ((d & 1) || ldx < d || (ldx & 1) || ldg < d || (ldg & 1) || ldk < m || (ldk & 1))
    return fail(EDRGP_ERR_ARG, "grad_tf32x3: d, ldx, ldg, ldk must be even and cover d / m");
  if (!aligned16(X) || !aligned16(Kfu) || !aligned16(G) || !aligned16(pack))
    return fail(EDRGP_ERR_ARG, "grad_tf32x3: X, Kfu, G and pack must be 16-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "grad_tf32x3: no CUDA device");
  cudaError_t e = edrgp::launch_grad_tf32(X, ldx, n, d, Kfu, ldk, sf2, ell, pack, m, G, ldg, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "grad_tf32x3");
}

size_t edrgp_pack_weights_tf32_bytes(int m) {
  if (m <= 0 || m > 512) return 0;
  return edrgp::pack_weights_tf32_bytes(m);
}

int edrgp_pack_weights_tf32(const double* M, int64_t ldm, double scale, int m, void* pack, void* stream) {
  if (!M || !pack || m <= 0 || ldm < m) return fail(EDRGP_ERR_ARG, "pack_weights_tf32: bad argument");
  if (m > 512) return fail(EDRGP_ERR_UNSUPPORTED, "pack_weights_tf32: m=%d > 512 is outside the TF32-split weights", m);
  if (!aligned16(pack)) return fail(EDRGP_ERR_ARG, "pack_weights_tf32: pack must be 16-byte aligned");
  cudaError_t e = edrgp::launch_pack_weights_tf32(M, ldm, scale, m, pack, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "pack_weights_tf32");
}

int edrgp_weights_tf32x3(const double* K, int64_t n, int m, int64_t ldk, const void* pack, const double* y,
                         const double* alpha, double c_ya, double* T, int64_t ldt, double* rowsum, void* stream) {
  if (!K || !pack || !T || n <= 0 || m <= 0) return fail(EDRGP_ERR_ARG, "weights_tf32x3: bad argument");
  if (m > 512) return fail(EDRGP_ERR_UNSUPPORTED, "weights_tf32x3: m=%d > 512 is outside the TF32-split weights", m);
  if (ldk < m || (ldk & 1) || ldt < m || (ldt & 1)) return fail(EDRGP_ERR_ARG, "weights_tf32x3: ldk / ldt must be even and >= m");
  if ((y == nullptr) != (alpha == nullptr)) return fail(EDRGP_ERR_ARG, "weights_tf32x3: y and alpha go together");
  if (!aligned16(K) || !aligned16(pack) || !aligned16(T)) return fail(EDRGP_ERR_ARG, "weights_tf32x3: K, pack and T must be 16-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "weights_tf32x3: no CUDA device");
  cudaError_t e = edrgp::launch_weights_tf32(K, n, m, ldk, pack, y, alpha, c_ya, T, ldt, rowsum, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "weights_tf32x3");
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

int edrgp_grad_gram_cached(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk, double sf2,
                            const double* pack, int m, double* G, int64_t ldg, double* C, void* workspace,
                            void* stream) {
  int rc = check_x("grad_gram_cached", X, n, d, pack, m);
  if (rc) return rc;
  if (ldx < d || (ldx & 1)) return fail(EDRGP_ERR_ARG, "grad_gram_cached: ldx must be even and >= d");
  if (G && (ldg < d || (ldg & 1))) return fail(EDRGP_ERR_ARG, "grad_gram_cached: ldg must be even and >= d");
  if (!Kfu || ldk < m || (ldk & 1) || !aligned16(Kfu))
    return fail(EDRGP_ERR_ARG, "grad_gram_cached: Kfu must be 16-byte aligned with an even leading dimension >= m");
  if (G && !aligned16(G)) return fail(EDRGP_ERR_ARG, "grad_gram_cached: G must be 16-byte aligned");
  if (C && !workspace) return fail(EDRGP_ERR_ARG, "grad_gram_cached: C needs a workspace");
  if (!edrgp::grad_gram_fused(d))
    return fail(EDRGP_ERR_UNSUPPORTED, "grad_gram_cached: needs d <= 64; use edrgp_grad_gram");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "grad_gram_cached: no CUDA device");
  cudaError_t e = edrgp::launch_grad_gram_cached(X, ldx, n, d, Kfu, ldk, sf2, pack, m, G, ldg, C, (double*)workspace,
                                                 sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "grad_gram_cached");
}

size_t edrgp_syrk_workspace_bytes(int64_t n, int k) {
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  return edrgp::gemm_tn_workspace_bytes(n, k, k, 1, sms);
}

static int check_tall(const char* who, const double* A, int64_t n, int k, int64_t lda) {
  if (!A || n <= 0 || k <= 0 || lda < k) return fail(EDRGP_ERR_ARG, "%s: bad argument", who);
  if ((lda & 1) || !aligned16(A)) return fail(EDRGP_ERR_ARG, "%s: leading dimension must be even, matrix 16-byte aligned", who);
  return EDRGP_OK;
}

int edrgp_syrk(const double* A, int64_t n, int k, int64_t lda, double* C, int64_t ldc, int accumulate,
               void* workspace, void* stream) {
  int rc = check_tall("syrk", A, n, k, lda);
  if (rc) return rc;
  if (!C || !workspace || ldc < k) return fail(EDRGP_ERR_ARG, "syrk: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "syrk: no CUDA device");
  cudaError_t e = edrgp::launch_gemm_tn(A, lda, k, nullptr, 0, 0, n, 1, nullptr, C, ldc, nullptr, accumulate,
                                        (double*)workspace, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "syrk");
}

int edrgp_inducing_stats(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double* P, int64_t ldp,
                         double* b_yy, int accumulate, void* workspace, void* stream) {
  int rc = check_tall("inducing_stats", Kfu, n, m, ldk);
  if (rc) return rc;
  if (!y || !P || !b_yy || !workspace || ldp < m) return fail(EDRGP_ERR_ARG, "inducing_stats: bad argument");
  if (!aligned16(y)) return fail(EDRGP_ERR_ARG, "inducing_stats: y must be 16-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "inducing_stats: no CUDA device");
  cudaError_t e = edrgp::launch_gemm_tn(Kfu, ldk, m, nullptr, 0, 0, n, 1, y, P, ldp, b_yy, accumulate,
                                        (double*)workspace, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "inducing_stats");
}

size_t edrgp_inducing_stats_i8_workspace_bytes(int64_t n, int m) {
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  if (n <= 0 || m <= 0 || m > 2048) return 0;
  return edrgp::inducing_stats_i8_workspace_bytes(n, m, sms);
}

int edrgp_inducing_stats_i8(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2, double* P,
                            int64_t ldp, double* b_yy, int accumulate, void* workspace, void* stream) {
  if (!Kfu || n <= 0 || m <= 0 || ldk < m || !P || ldp < m || !workspace) return fail(EDRGP_ERR_ARG, "inducing_stats_i8: bad argument");
  if (m > 2048) return fail(EDRGP_ERR_UNSUPPORTED, "inducing_stats_i8: m=%d > 2048", m);
  if ((y == nullptr) != (b_yy == nullptr)) return fail(EDRGP_ERR_ARG, "inducing_stats_i8: y and b_yy go together");
  if (!(sf2 > 0.0) || !(sf2 < 1e150)) return fail(EDRGP_ERR_ARG, "inducing_stats_i8: the kernel variance must be positive and finite");
  if (reinterpret_cast<uintptr_t>(workspace) & 127) return fail(EDRGP_ERR_ARG, "inducing_stats_i8: the workspace must be 128-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "inducing_stats_i8: no CUDA device");
  cudaError_t e = edrgp::launch_inducing_stats_i8(Kfu, n, m, ldk, y, sf2, P, ldp, b_yy, accumulate, workspace, sms,
                                                  (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "inducing_stats_i8");
}

size_t edrgp_gemm_tn_workspace_bytes(int64_t n, int ka, int kb) {
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  return edrgp::gemm_tn_workspace_bytes(n, ka, kb, 0, sms);
}

int edrgp_gemm_tn(const double* A, int64_t lda, int ka, const double* B, int64_t ldb, int kb, int64_t n, double* C,
                  int64_t ldc, int accumulate, void* workspace, void* stream) {
  int rc = check_tall("gemm_tn", A, n, ka, lda);
  if (rc) return rc;
  rc = check_tall("gemm_tn", B, n, kb, ldb);
  if (rc) return rc;
  if (!C || !workspace || ldc < kb) return fail(EDRGP_ERR_ARG, "gemm_tn: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "gemm_tn: no CUDA device");
  cudaError_t e = edrgp::launch_gemm_tn(A, lda, ka, B, ldb, kb, n, 0, nullptr, C, ldc, nullptr, accumulate,
                                        (double*)workspace, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "gemm_tn");
}

size_t edrgp_weights_workspace_bytes(int64_t n, int m) { return edrgp::weights_workspace_bytes(n, m); }

int edrgp_weights(const double* Kfu, int64_t n, int m, int64_t ldk, const double* M, int64_t ldm, const double* y,
                  const double* alpha, double c_ya, double c_km, double* T, int64_t ldt, double* rowsum,
                  double* colsum, int accumulate, void* workspace, void* stream) {
  int rc = check_tall("weights", Kfu, n, m, ldk);
  if (rc) return rc;
  if (!M || ldm < m || (ldm & 1) || !aligned16(M)) return fail(EDRGP_ERR_ARG, "weights: M needs an even leading dimension >= m and 16-byte alignment");
  if (T && (ldt < m || (ldt & 1) || !aligned16(T))) return fail(EDRGP_ERR_ARG, "weights: T needs an even leading dimension >= m and 16-byte alignment");
  if ((y == nullptr) != (alpha == nullptr)) return fail(EDRGP_ERR_ARG, "weights: y and alpha go together");
  if ((rowsum || colsum) && !workspace) return fail(EDRGP_ERR_ARG, "weights: sums need a workspace");
  cudaError_t e = edrgp::launch_weights(Kfu, n, m, ldk, M, ldm, y, alpha, c_ya, c_km, T, ldt, rowsum, colsum, accumulate,
                                        (double*)workspace, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "weights");
}

int edrgp_kmm(const double* Zp, int64_t ldz, const double* pack, int m, int d, double sf2, double jitter,
              double* Kmm, int64_t ldk, int multiply, int finish, void* stream) {
  int rc = check_x("kmm", Zp, m, d, pack, m);
  if (rc) return rc;
  if (ldz < d || (ldz & 1)) return fail(EDRGP_ERR_ARG, "kmm: ldz must be even and >= d");
  if (!Kmm || ldk < m || (ldk & 1) || !aligned16(Kmm))
    return fail(EDRGP_ERR_ARG, "kmm: Kmm must be 16-byte aligned with an even leading dimension >= m");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "kmm: no CUDA device");
  cudaError_t e = edrgp::launch_kuf(Zp, ldz, m, d, pack, m, sf2, Kmm, ldk, multiply, nullptr, nullptr, nullptr, sms,
                                    (cudaStream_t)stream);
  if (e == cudaSuccess && finish) e = edrgp::launch_kmm_fix(Kmm, m, ldk, sf2, jitter, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "kmm");
}

size_t edrgp_solve_workspace_bytes(int m) { return (size_t)2 * m * m * sizeof(double); }

int edrgp_solve(double* Kmm, const double* P, const double* b, int m, double beta, double* LB, double* alpha,
                double* c, double* scalars, int* info, void* workspace, void* stream) {
  if (!Kmm || !P || !b || !LB || !alpha || !c || !scalars || !info || !workspace || m <= 0)
    return fail(EDRGP_ERR_ARG, "solve: bad argument");
  cudaError_t e = edrgp::launch_solve(Kmm, P, b, m, beta, LB, alpha, c, scalars, info, (double*)workspace,
                                      (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "solve");
}

size_t edrgp_vfe_grad_small_workspace_bytes(int m) { return m > 0 ? (size_t)3 * m * m * sizeof(double) : 0; }

int edrgp_vfe_grad_small(const double* LB, const double* Lm, const double* B, const double* c, int m, double beta,
                         double* Msym, int64_t ldm, double* Dsym, double* sumAE, void* workspace, void* stream) {
  if (!LB || !Lm || !B || !c || !Msym || !Dsym || !sumAE || !workspace || m <= 0 || ldm < m)
    return fail(EDRGP_ERR_ARG, "vfe_grad_small: bad argument");
  cudaError_t e = edrgp::launch_vfe_grad_small(LB, Lm, B, c, m, beta, Msym, ldm, Dsym, sumAE, (double*)workspace,
                                               (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "vfe_grad_small");
}

int edrgp_potrf(double* A, int m, int64_t ld, int* info, void* stream) {
  if (!A || !info || m <= 0 || ld < m) return fail(EDRGP_ERR_ARG, "potrf: bad argument");
  cudaError_t e = edrgp::launch_potrf(A, m, ld, info, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "potrf");
}

int edrgp_posv(double* A, int m, int64_t ld, double* L, int64_t ldl, double* rhs, double* x, int* info, void* stream) {
  if (!A || !L || !info || m <= 0 || ld < m || ldl < m || (rhs && (!x || x == rhs)) || A == L)
    return fail(EDRGP_ERR_ARG, "posv: bad argument");
  cudaError_t e = edrgp::launch_posv(A, m, ld, L, ldl, rhs, x, info, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "posv");
}

int edrgp_trsm(const double* L, int m, double* B, int nrhs, int trans, void* stream) {
  if (!L || !B || m <= 0 || nrhs <= 0) return fail(EDRGP_ERR_ARG, "trsm: bad argument");
  cudaError_t e = edrgp::launch_trsm(L, m, m, B, nrhs, nrhs, trans, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "trsm");
}

size_t edrgp_eigh_workspace_bytes(int d) { return d > 0 ? edrgp::eigh_workspace_doubles(d) * sizeof(double) : 0; }

int edrgp_eigh(double* C, int d, double* evals, double* comps, int* sweeps, void* workspace, void* stream) {
  if (!C || !evals || !comps || !workspace || d <= 0) return fail(EDRGP_ERR_ARG, "eigh: bad argument");
  cudaError_t e = edrgp::launch_eigh(C, d, (double*)workspace, evals, comps, sweeps, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "eigh");
}

size_t edrgp_col_moments_workspace_bytes(int d) {
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  return edrgp::col_moments_workspace_bytes(d, sms);
}

int edrgp_col_moments(const double* X, int64_t n, int d, const double* shift, const double* weight, double* out,
                      int accumulate, void* workspace, void* stream) {
  if (!X || !out || !workspace || n <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "col_moments: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "col_moments: no CUDA device");
  cudaError_t e = edrgp::launch_col_moments(X, n, d, shift, weight, out, accumulate, (double*)workspace, sms,
                                            (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "col_moments");
}

int edrgp_count_nonfinite(const double* X, int64_t total, unsigned int* count, void* stream) {
  if (!X || !count || total < 0) return fail(EDRGP_ERR_ARG, "count_nonfinite: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "count_nonfinite: no CUDA device");
  if (total == 0) return EDRGP_OK;
  cudaError_t e = edrgp::launch_count_nonfinite(X, total, count, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "count_nonfinite");
}

int edrgp_standardize(const double* X, int64_t n, int d, const double* mean, const double* scale, double* out,
                      void* stream) {
  if (!X || !mean || !scale || !out || n <= 0 || d <= 0) return fail(EDRGP_ERR_ARG, "standardize: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "standardize: no CUDA device");
  cudaError_t e = edrgp::launch_standardize(X, n, d, mean, scale, out, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "standardize");
}

int edrgp_project_dmma(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int k, double* out,
                       int64_t ldo, void* stream) {
  int rc = check_x("project_dmma", X, n, d, pack, k);
  if (rc) return rc;
  if (ldx < d || (ldx & 1)) return fail(EDRGP_ERR_ARG, "project_dmma: ldx must be even and >= d");
  if (!out || ldo < k || (ldo & 1) || !aligned16(out))
    return fail(EDRGP_ERR_ARG, "project_dmma: out must be 16-byte aligned with an even leading dimension >= k");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "project_dmma: no CUDA device");
  cudaError_t e = edrgp::launch_kuf(X, ldx, n, d, pack, k, 1.0, out, ldo, 0, nullptr, nullptr, nullptr, sms,
                                    (cudaStream_t)stream, 1);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "project_dmma");
}

int edrgp_project(const double* X, int64_t n, int d, const double* V, int k, double* out, void* stream) {
  if (!X || !V || !out || n <= 0 || d <= 0 || k <= 0) return fail(EDRGP_ERR_ARG, "project: bad argument");
  const int sms = sm_count_cached();
  if (sms <= 0) return fail(EDRGP_ERR_CUDA, "project: no CUDA device");
  cudaError_t e = edrgp::launch_project(X, n, d, V, k, out, sms, (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "project");
}


// ---------------------------------------------------------------------------------------------------------------
// Fixed-hyper-parameter sweep composites (sweep.cu): everything between two collectives in one call
// ---------------------------------------------------------------------------------------------------------------
namespace {
// statistics mode of the composite sweep: 0 = FP64 DMMA, 1 = exact-product INT8 slices.  Process-wide, because the
// workspace layout depends on it; the default comes from EDRGP_STATS (fp64 | int8x6).
int g_stats_mode = -1;
int stats_mode() {
  if (g_stats_mode < 0) {
    const char* s = getenv("EDRGP_STATS");
    g_stats_mode = (s && (!strcmp(s, "int8x6") || !strcmp(s, "int8") || !strcmp(s, "1"))) ? 1 : 0;
  }
  return g_stats_mode;
}
// the INT8 route covers what its kernel covers; outside that the FP64 reduction runs (same results to 1e-13)
int stats_mode_for(int m, double sf2) { return (stats_mode() == 1 && m <= 2048 && sf2 < 1e150) ? 1 : 0; }

// NVLink peer exchange (peer.cu): a workspace that has been bound to the ranks' exchange buffers runs its three
// reductions inside the composite calls; epochs count the uses of each payload (all ranks call in lockstep).
struct PeerBinding {
  edrgp::PeerCtx ctx;
  int m, d;
  int* epoch;        // [PEER_COLLECTIVES], owned by the BUFFER set (several workspaces may be bound to the same buffers)
};
struct PeerEpochs { int v[edrgp::PEER_COLLECTIVES]; };
std::mutex g_peer_mu;
std::unordered_map<void*, PeerBinding> g_peer_bindings;
std::unordered_map<void*, PeerEpochs> g_peer_epochs;       // keyed by the rank's own exchange buffer
PeerBinding* peer_binding(void* workspace, int m, int d, int world) {
  std::lock_guard<std::mutex> lk(g_peer_mu);
  auto it = g_peer_bindings.find(workspace);
  if (it == g_peer_bindings.end()) return nullptr;
  PeerBinding* b = &it->second;
  return (b->m == m && b->d == d && b->ctx.world == world) ? b : nullptr;
}

struct FixedCtx {
  int64_t off[edrgp::FS_NREGIONS];
  int sms;
  double* ws;
  double* at(int region) const { return ws + off[region]; }
  unsigned int* flag() const { return reinterpret_cast<unsigned int*>(at(edrgp::FS_TAIL)); }
  int* info() const { return reinterpret_cast<int*>(at(edrgp::FS_TAIL)) + 1; }
};
int fixed_ctx(const char* who, int64_t n, int d, int m, int64_t chunk_rows, int world, void* ws, FixedCtx* c) {
  if (n <= 0 || d <= 0 || m <= 0 || chunk_rows <= 0 || world <= 0 || !ws) return fail(EDRGP_ERR_ARG, "%s: bad argument", who);
  if ((d & 1) || d > 64) return fail(EDRGP_ERR_UNSUPPORTED, "%s: the composite sweep covers even d <= 64 (got %d)", who, d);
  if (chunk_rows & 1) return fail(EDRGP_ERR_ARG, "%s: chunk_rows must be even", who);
  if (!aligned16(ws)) return fail(EDRGP_ERR_ARG, "%s: the workspace must be 16-byte aligned", who);
  c->sms = sm_count_cached();
  if (c->sms <= 0) return fail(EDRGP_ERR_CUDA, "%s: no CUDA device", who);
  edrgp::fixed_layout(n, d, m, chunk_rows, world, c->sms, c->off, stats_mode() == 1 && m <= 2048);
  c->ws = (double*)ws;
  return EDRGP_OK;
}
}  // namespace

size_t edrgp_fixed_layout(int64_t n, int d, int m, int64_t chunk_rows, int world, int64_t* offsets) {
  int sms = sm_count_cached();
  if (sms <= 0) sms = 160;
  if (n <= 0 || d <= 0 || m <= 0 || chunk_rows <= 0 || world <= 0 || !offsets) return 0;
  return edrgp::fixed_layout(n, d, m, chunk_rows, world, sms, offsets, stats_mode() == 1 && m <= 2048) * sizeof(double);
}

int edrgp_set_stats_mode(int mode) {
  if (mode != 0 && mode != 1) return fail(EDRGP_ERR_ARG, "set_stats_mode: 0 (fp64) or 1 (int8x6)");
  g_stats_mode = mode;
  return EDRGP_OK;
}

int edrgp_get_stats_mode(void) { return stats_mode(); }

int edrgp_fixed_begin(const double* X, int64_t ldx, int64_t n, int d, const double* y, const double* Z, int64_t ldz,
                      const double* ell, int m, double sf2, int64_t chunk_rows, double* Kfu, int64_t ldk, int rank,
                      int world, void* h2d, int64_t h2d_ahead, void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_begin", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  if (!X || !y || !Z || !ell || !Kfu || ldx < d || (ldx & 1) || ldz != d || ldk < m || (ldk & 1) || rank < 0 || rank >= world)
    return fail(EDRGP_ERR_ARG, "fixed_begin: bad argument");
  if (!aligned16(X) || !aligned16(Kfu) || !aligned16(y)) return fail(EDRGP_ERR_ARG, "fixed_begin: X, y and Kfu must be 16-byte aligned");
  if (!(sf2 > 0.0) || !(sf2 < 1e300)) return fail(EDRGP_ERR_ARG, "fixed_begin: the kernel variance must be positive and finite");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  double* scratch = c.at(edrgp::FS_SCRATCH);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 4 * c.sms);
  // only the flag / info word: N, mean and std of an earlier normalisation stay valid when this call runs with
  // targets that are already normalised
  if ((e = cudaMemsetAsync(c.at(edrgp::FS_TAIL), 0, sizeof(double), st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  if ((e = cudaMemsetAsync(c.at(edrgp::FS_TABLE), 0, (size_t)4 * world * sizeof(double), st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  if ((e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  if ((e = edrgp::launch_pack(Z, ell, nullptr, 1.0, m, d, c.at(edrgp::FS_PACK_K), st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  const int64_t rows = chunk_rows < n ? chunk_rows : n;
  // the first cross-covariance block does not need the targets: it goes first, so that the device is busy while
  // the host walks through the rest of this call and (multi-rank) the gather of the moments table
  // rows (and targets) still on their way from the host: the stream waits for the blocks this call touches
  if (h2d && (e = edrgp::h2d_wait((edrgp::H2DTransfer*)h2d, rows, h2d_ahead, st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  {
    StageScope t(EDRGP_STAGE_KUF, st);
    if ((e = edrgp::launch_kuf(X, ldx, rows, d, c.at(edrgp::FS_PACK_K), m, sf2, Kfu, ldk, 0, nullptr, nullptr, nullptr, c.sms,
                               st, 0, c.flag())) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  }
  if (h2d && (e = edrgp::h2d_wait_side((edrgp::H2DTransfer*)h2d, st)) != cudaSuccess) return cuda_fail(e, "fixed_begin");
  StageScope t(EDRGP_STAGE_TARGETS, st);
  if ((e = edrgp::launch_target_moments(y, n, scratch, ticket, c.at(edrgp::FS_TABLE) + 4 * rank, c.sms, st)) != cudaSuccess)
    return cuda_fail(e, "fixed_begin");
  if (PeerBinding* pb = peer_binding(workspace, m, d, world)) {
    // peer exchange: this rank's row goes into every rank's table, then its flag (no collective call follows)
    if (pb->ctx.rank != rank) return fail(EDRGP_ERR_ARG, "fixed_begin: the workspace is bound as another rank");
    const int ep = ++pb->epoch[edrgp::PEER_COLL_TABLE];
    if ((e = edrgp::launch_peer_push_table(pb->ctx, c.at(edrgp::FS_TABLE) + 4 * rank, ep, st)) != cudaSuccess)
      return cuda_fail(e, "fixed_begin");
  }
  return EDRGP_OK;
}

int edrgp_fixed_stats(const double* X, int64_t ldx, int64_t n, int d, const double* y, int m, double sf2,
                      int64_t chunk_rows, double* Kfu, int64_t ldk, int normalize, int world, void* h2d, int64_t h2d_ahead,
                      void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_stats", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  if (!X || !y || !Kfu || ldx < d || (ldx & 1) || ldk < m || (ldk & 1)) return fail(EDRGP_ERR_ARG, "fixed_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  double* yt = c.at(edrgp::FS_YT);
  PeerBinding* pb = peer_binding(workspace, m, d, world);
  if (pb) {
    StageScope t(EDRGP_STAGE_TARGETS, st);
    if ((e = edrgp::launch_peer_table_wait(pb->ctx, pb->epoch[edrgp::PEER_COLL_TABLE], c.at(edrgp::FS_TABLE), c.flag(), st)) !=
        cudaSuccess) return cuda_fail(e, "fixed_stats");
  }
  {
    StageScope t(EDRGP_STAGE_TARGETS, st);
    e = edrgp::launch_target_standardize(c.at(edrgp::FS_TABLE), world, y, n, yt, c.at(edrgp::FS_TAIL) + 1, normalize, c.flag(),
                                            c.sms, st);
  }
  if (e != cudaSuccess) return cuda_fail(e, "fixed_stats");
  const double* targets = normalize ? yt : y;
  // peer exchange: the rank's partial statistics are written straight into its exchange buffer (the peers read them
  // from there while they assemble their systems); the workspace region receives the SUM in edrgp_fixed_posterior
  const int ep_stats = pb ? ++pb->epoch[edrgp::PEER_COLL_STATS] : 0;
  double* P = pb ? edrgp::peer_partial(pb->ctx, edrgp::PEER_STATS, ep_stats, (size_t)m * m + m + 1) : c.at(edrgp::FS_STATS);
  double* byy = P + (size_t)m * m;
  const int i8 = stats_mode_for(m, sf2);
  void* i8ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(c.at(edrgp::FS_SCRATCH)) + 1023) & ~(uintptr_t)1023);
  for (int64_t s = 0; s < n; s += chunk_rows) {
    const int64_t rows = n - s < chunk_rows ? n - s : chunk_rows;
    double* Kc = Kfu + s * ldk;
    if (s > 0 && h2d && (e = edrgp::h2d_wait((edrgp::H2DTransfer*)h2d, s + rows, h2d_ahead, st)) != cudaSuccess)
      return cuda_fail(e, "fixed_stats");
    if (s > 0) {
      StageScope t(EDRGP_STAGE_KUF, st);
      if ((e = edrgp::launch_kuf(X + s * ldx, ldx, rows, d, c.at(edrgp::FS_PACK_K), m, sf2, Kc, ldk, 0, nullptr, nullptr,
                                 nullptr, c.sms, st, 0, c.flag())) != cudaSuccess) return cuda_fail(e, "fixed_stats");
    }
    StageScope t(EDRGP_STAGE_STATS, st);
    if (i8) {
      if ((e = edrgp::launch_i8_block(Kc, rows, m, ldk, targets + s, sf2, s == 0, i8ws, c.sms, st)) != cudaSuccess)
        return cuda_fail(e, "fixed_stats");
      if (s + rows >= n &&
          (e = edrgp::launch_i8_finish(m, sf2, 1, P, m, byy, 0, i8ws, c.sms, st)) != cudaSuccess)
        return cuda_fail(e, "fixed_stats");
    } else if ((e = edrgp::launch_gemm_tn(Kc, ldk, m, nullptr, 0, 0, rows, 1, targets + s, P, m, byy, s > 0,
                                          c.at(edrgp::FS_SCRATCH), c.sms, st)) != cudaSuccess) {
      return cuda_fail(e, "fixed_stats");
    }
  }
  if (pb && (e = edrgp::launch_peer_signal(pb->ctx, edrgp::PEER_COLL_STATS, ep_stats, st)) != cudaSuccess)
    return cuda_fail(e, "fixed_stats");
  return EDRGP_OK;
}

int edrgp_fixed_posterior(const double* Z, int64_t ldz, int64_t n, int d, int m, double sf2, double jitter, double beta,
                          int64_t chunk_rows, int world, void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_posterior", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  if (!Z || ldz != d || !(beta > 0.0)) return fail(EDRGP_ERR_ARG, "fixed_posterior: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  const int64_t lds = m + (m & 1);
  StageScope t(EDRGP_STAGE_SOLVE, st);
  double* S = c.at(edrgp::FS_S);
  const double* P = c.at(edrgp::FS_STATS);
  if ((e = edrgp::launch_kuf(Z, ldz, m, d, c.at(edrgp::FS_PACK_K), m, sf2, S, lds, 0, nullptr, nullptr, nullptr, c.sms, st)) !=
      cudaSuccess) return cuda_fail(e, "fixed_posterior");
  if (PeerBinding* pb = peer_binding(workspace, m, d, world)) {
    // the reduction over ranks happens while the system is assembled: partials read from the peers over NVLink
    if ((e = edrgp::launch_form_system_peer(pb->ctx, pb->epoch[edrgp::PEER_COLL_STATS], S, m, lds, sf2, jitter, beta,
                                            c.at(edrgp::FS_STATS), c.at(edrgp::FS_RHS), c.flag(), st)) != cudaSuccess)
      return cuda_fail(e, "fixed_posterior");
  } else if ((e = edrgp::launch_form_system(S, m, lds, sf2, jitter, beta, P, m, P + (size_t)m * m, c.at(edrgp::FS_RHS), st)) !=
      cudaSuccess) return cuda_fail(e, "fixed_posterior");
  if ((e = edrgp::launch_posv(S, m, lds, c.at(edrgp::FS_L), m, c.at(edrgp::FS_RHS), c.at(edrgp::FS_ALPHA), c.info(), st)) !=
      cudaSuccess) return cuda_fail(e, "fixed_posterior");
  return EDRGP_OK;
}

int edrgp_fixed_grad(const double* X, int64_t ldx, int64_t n, int d, const double* Kfu, int64_t ldk, const double* Z,
                     int64_t ldz, const double* ell, int m, double sf2, double coef_scale, const double* dev_scale,
                     double* G, int64_t ldg, int64_t chunk_rows, int world, void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_grad", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  if (!X || !Kfu || !Z || !ell || ldx < d || (ldx & 1) || ldz != d || ldk < m || (ldk & 1) || (G && (ldg < d || (ldg & 1))))
    return fail(EDRGP_ERR_ARG, "fixed_grad: bad argument");
  if (!aligned16(X) || !aligned16(Kfu) || (G && !aligned16(G))) return fail(EDRGP_ERR_ARG, "fixed_grad: X, Kfu and G must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  double* res = c.at(edrgp::FS_RESULT);
  double* C = res + d + (size_t)d * d;
  StageScope t(EDRGP_STAGE_GRAD, st);
  if ((e = edrgp::launch_pack(Z, ell, c.at(edrgp::FS_ALPHA), coef_scale, m, d, c.at(edrgp::FS_PACK_G), st, dev_scale)) !=
      cudaSuccess) return cuda_fail(e, "fixed_grad");
  if ((e = edrgp::launch_grad_gram_cached(X, ldx, n, d, Kfu, ldk, sf2, c.at(edrgp::FS_PACK_G), m, G, ldg, C,
                                          c.at(edrgp::FS_SCRATCH), c.sms, st)) != cudaSuccess) return cuda_fail(e, "fixed_grad");
  if ((e = cudaMemcpyAsync(C + (size_t)d * d, c.at(edrgp::FS_TAIL), 4 * sizeof(double), cudaMemcpyDeviceToDevice, st)) !=
      cudaSuccess) return cuda_fail(e, "fixed_grad");
  return EDRGP_OK;
}

int edrgp_fixed_reduce_gram(int64_t n, int d, int m, int64_t chunk_rows, int world, void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_reduce_gram", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  PeerBinding* pb = peer_binding(workspace, m, d, world);
  if (!pb) return fail(EDRGP_ERR_UNSUPPORTED, "fixed_reduce_gram: the workspace is not bound to peer exchange buffers");
  cudaStream_t st = (cudaStream_t)stream;
  double* res = c.at(edrgp::FS_RESULT);
  double* C = res + d + (size_t)d * d;
  const int ep = ++pb->epoch[edrgp::PEER_COLL_GRAM];
  double* part = edrgp::peer_partial(pb->ctx, edrgp::PEER_GRAM, ep, (size_t)d * d);
  StageScope t(EDRGP_STAGE_GRAD, st);
  cudaError_t e;
  if ((e = cudaMemcpyAsync(part, C, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
    return cuda_fail(e, "fixed_reduce_gram");
  if ((e = edrgp::launch_peer_signal(pb->ctx, edrgp::PEER_COLL_GRAM, ep, st)) != cudaSuccess) return cuda_fail(e, "fixed_reduce_gram");
  if ((e = edrgp::launch_reduce_gram_peer(pb->ctx, ep, d * d, C, c.flag(), st)) != cudaSuccess)
    return cuda_fail(e, "fixed_reduce_gram");
  // (the tail copy behind C was made by edrgp_fixed_grad; a time-out raised here reaches the host through the flag word)
  if ((e = cudaMemcpyAsync(C + (size_t)d * d, c.at(edrgp::FS_TAIL), 4 * sizeof(double), cudaMemcpyDeviceToDevice, st)) !=
      cudaSuccess) return cuda_fail(e, "fixed_reduce_gram");
  return EDRGP_OK;
}

size_t edrgp_peer_layout(int m, int d, int world, int64_t* offsets) {
  if (m <= 0 || d <= 0 || world <= 0 || world > edrgp::PEER_MAX_WORLD || !offsets) return 0;
  return edrgp::peer_layout(m, d, world, offsets) * sizeof(double);
}

int edrgp_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle) {
  if (!bytes || !dev_ptr || !ipc_handle) return fail(EDRGP_ERR_ARG, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == EDRGP_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "peer_alloc");
  if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess) { cudaFree(p); return cuda_fail(e, "peer_alloc"); }
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) { cudaFree(p); return cuda_fail(e, "peer_alloc"); }
  memcpy(ipc_handle, &h, sizeof(h));
  *dev_ptr = p;
  {
    std::lock_guard<std::mutex> lk(g_peer_mu);
    g_peer_epochs.erase(p);           // a fresh buffer (flags zeroed) starts at epoch 0 even if the address is an old one
  }
  return EDRGP_OK;
}

int edrgp_peer_open(const void* ipc_handle, void** dev_ptr) {
  if (!ipc_handle || !dev_ptr) return fail(EDRGP_ERR_ARG, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "peer_open");
  *dev_ptr = p;
  return EDRGP_OK;
}

int edrgp_peer_close(void* dev_ptr) {
  if (!dev_ptr) return EDRGP_OK;
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "peer_close");
}

int edrgp_peer_free(void* dev_ptr) {
  if (!dev_ptr) return EDRGP_OK;
  {
    std::lock_guard<std::mutex> lk(g_peer_mu);
    for (auto it = g_peer_bindings.begin(); it != g_peer_bindings.end();)
      it = it->second.ctx.base[it->second.ctx.rank] == dev_ptr ? g_peer_bindings.erase(it) : std::next(it);
    g_peer_epochs.erase(dev_ptr);
  }
  cudaError_t e = cudaFree(dev_ptr);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "peer_free");
}

int edrgp_fixed_bind_peers(void* workspace, void* const* bases, int rank, int world, int m, int d) {
  if (!workspace) return fail(EDRGP_ERR_ARG, "fixed_bind_peers: bad argument");
  std::lock_guard<std::mutex> lk(g_peer_mu);
  if (!bases) { g_peer_bindings.erase(workspace); return EDRGP_OK; }
  if (world < 1 || world > edrgp::PEER_MAX_WORLD || rank < 0 || rank >= world || m <= 0 || d <= 0)
    return fail(EDRGP_ERR_ARG, "fixed_bind_peers: 1 <= world <= %d, 0 <= rank < world", edrgp::PEER_MAX_WORLD);
  PeerBinding b{};
  for (int r = 0; r < world; ++r) {
    if (!bases[r]) return fail(EDRGP_ERR_ARG, "fixed_bind_peers: null exchange buffer of rank %d", r);
    b.ctx.base[r] = (double*)bases[r];
  }
  edrgp::peer_layout(m, d, world, b.ctx.off);
  b.ctx.rank = rank; b.ctx.world = world; b.m = m; b.d = d;
  b.epoch = g_peer_epochs[bases[rank]].v;                 // (value-initialised to 0 on first use; node addresses are stable)
  g_peer_bindings[workspace] = b;
  return EDRGP_OK;
}

int edrgp_fixed_eigh(int64_t n, int d, int m, int64_t chunk_rows, int world, void* workspace, void* stream) {
  FixedCtx c;
  int rc = fixed_ctx("fixed_eigh", n, d, m, chunk_rows, world, workspace, &c);
  if (rc) return rc;
  double* res = c.at(edrgp::FS_RESULT);
  StageScope t(EDRGP_STAGE_EIGH, (cudaStream_t)stream);
  cudaError_t e = edrgp::launch_eigh(res + d + (size_t)d * d, d, c.at(edrgp::FS_SCRATCH), res, res + d, nullptr,
                                     (cudaStream_t)stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "fixed_eigh");
}

void* edrgp_h2d_open(const void* host, void* dev, int64_t rows, size_t row_bytes, size_t dst_pitch, int64_t block_rows,
                     int threads, int slots, void* order_after_stream, const void* side_host, void* side_dev,
                     size_t side_bytes) {
  if (!host || !dev || rows <= 0 || row_bytes == 0 || dst_pitch < row_bytes || block_rows <= 0) {
    fail(EDRGP_ERR_ARG, "h2d_open: bad argument");
    return nullptr;
  }
  cudaError_t e = cudaSuccess;
  edrgp::H2DTransfer* t = edrgp::h2d_open(host, dev, rows, row_bytes, dst_pitch, block_rows, threads, slots,
                                          (cudaStream_t)order_after_stream, side_host, side_dev, side_bytes, &e);
  if (!t) cuda_fail(e, "h2d_open");
  return t;
}

int edrgp_h2d_wait(void* handle, int64_t upto_row, int64_t ahead_rows, void* consumer_stream) {
  if (!handle) return fail(EDRGP_ERR_ARG, "h2d_wait: bad argument");
  cudaError_t e = edrgp::h2d_wait((edrgp::H2DTransfer*)handle, upto_row, ahead_rows, (cudaStream_t)consumer_stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "h2d_wait");
}

int edrgp_h2d_wait_side(void* handle, void* consumer_stream) {
  if (!handle) return fail(EDRGP_ERR_ARG, "h2d_wait_side: bad argument");
  cudaError_t e = edrgp::h2d_wait_side((edrgp::H2DTransfer*)handle, (cudaStream_t)consumer_stream);
  return e == cudaSuccess ? EDRGP_OK : cuda_fail(e, "h2d_wait_side");
}

int edrgp_h2d_staged(void* handle) { return handle && edrgp::h2d_staged((const edrgp::H2DTransfer*)handle) ? 1 : 0; }

int edrgp_h2d_close(void* handle) {
  edrgp::h2d_close((edrgp::H2DTransfer*)handle);
  return EDRGP_OK;
}

int edrgp_timing_begin(void) {
  std::lock_guard<std::mutex> lk(g_timer_mu);
  g_timer.spans.clear();
  g_timer.used = 0;
  g_timer.on = true;
  return EDRGP_OK;
}

int edrgp_timing_end(double* ms, int* count) {
  std::lock_guard<std::mutex> lk(g_timer_mu);
  g_timer.on = false;
  for (int i = 0; i < EDRGP_STAGE_COUNT; ++i) { if (ms) ms[i] = 0.0; if (count) count[i] = 0; }
  for (const auto& sp : g_timer.spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    if (e != cudaSuccess) return cuda_fail(e, "timing_end");
    float t = 0.f;
    if ((e = cudaEventElapsedTime(&t, sp.a, sp.b)) != cudaSuccess) return cuda_fail(e, "timing_end");
    if (ms) ms[sp.stage] += (double)t;
    if (count) count[sp.stage] += 1;
  }
  g_timer.spans.clear();
  g_timer.used = 0;
  return EDRGP_OK;
}

}  // extern "C"
