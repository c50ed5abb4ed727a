// The fixed-hyper-parameter EDR sweep as a handful of composite calls (declared in include/edrgp_b200.h as
// edrgp_fixed_*): everything between two collectives is enqueued by ONE call, so that a rank whose row shard takes
// a few milliseconds is not waiting for its host between kernels.
//
//   edrgp_fixed_begin      pack(Z, l)  ->  Kfu block 0  ->  target moments of this rank            [all-gather table]
//   edrgp_fixed_stats      global mean / std(y), standardised targets, P = Kfu^T Kfu, b = Kfu^T y, y^T y over all
//                          row blocks (the cross-covariance of blocks 1.. is computed here)        [all-reduce stats]
//   edrgp_fixed_posterior  S = Kuu + jitter I + beta P,  alpha = S^-1 beta b (edrgp_posv)
//   edrgp_fixed_grad       pack(Z, l, alpha std(y))  ->  posterior-mean gradients from the stored Kfu, C = G^T G
//                                                                                                    [all-reduce C]
//   edrgp_fixed_eigh       eigh(C) next to C and the deferred-check words: one read-back for the host
//
// Reference path: one `estimator.fit` + `predict_gradient` + `SVDTransformer.fit` of edrgp/base.py:435-466 at fixed
// hyper-parameters (edrgp/gp_model/base.py:46-70,208-222; edrgp/utils.py:123-157).  This file holds the layout of the
// workspace the calls share and the three small kernels that are new here; the n-scale kernels are the library's own
// (pipeline.cu, syrk.cu) and the solvers linalg.cu's.
#include "common.cuh"
#include "launch.h"
#include "sweep.h"

namespace edrgp {

// ---------------------------------------------------------------------------------------------------------------
// Target normaliser (GPy Standardize: mean and population std over ALL rows of all ranks) with ONE collective:
// every rank reduces [n_r, c_r, S1_r, S2_r] with S_k = sum (y - c_r)^k about its own pivot c_r = y[0] in one pass;
// after the table of all ranks has been gathered,  mean = sum (n_r c_r + S1_r) / N  and
//   sum (y - mean)^2 = sum_r [ S2_r + 2 (c_r - mean) S1_r + n_r (c_r - mean)^2 ]      (exact identity),
// which has no cancellation as long as the pivot lies within a few std of the mean (it is a sample of y).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TM_THREADS = 256;

__global__ void __launch_bounds__(TM_THREADS) target_moments_kernel(const double* __restrict__ y, int64_t n,
                                                                    double* __restrict__ part, unsigned int* ticket,
                                                                    double* __restrict__ slot) {
  __shared__ double red[2][TM_THREADS / 32];
  __shared__ int last;
  const double c = y[0];
  double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
  const int64_t stride = (int64_t)gridDim.x * TM_THREADS;
  int64_t i = (int64_t)blockIdx.x * TM_THREADS + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {
    const double u = y[i] - c, v = y[i + stride] - c;
    s1a += u; s2a = fma(u, u, s2a);
    s1b += v; s2b = fma(v, v, s2b);
  }
  if (i < n) { const double u = y[i] - c; s1a += u; s2a = fma(u, u, s2a); }
  double s1 = s1a + s1b, s2 = s2a + s2b;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < TM_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA sums the partials in block order (deterministic for a given grid)
  if (warp == 0) {
    double a = 0.0, b = 0.0;
    for (int k = lane; k < (int)gridDim.x; k += 32) { a += part[2 * k]; b += part[2 * k + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { slot[0] = (double)n; slot[1] = c; slot[2] = a; slot[3] = b; *ticket = 0u; }
  }
}

// tail3 = [N, mean, std] from the gathered table; yt = (y - mean) / std (normalize) -- every CTA recomputes the
// three scalars from the (tiny) table instead of waiting for one that does.
__global__ void __launch_bounds__(256) target_standardize_kernel(const double* __restrict__ table, int world,
                                                                 const double* __restrict__ y, int64_t n,
                                                                 double* __restrict__ yt, double* __restrict__ tail3,
                                                                 int normalize, unsigned int* __restrict__ flag) {
  double N = 0.0, s = 0.0;
  for (int r = 0; r < world; ++r) { N += table[4 * r]; s += fma(table[4 * r], table[4 * r + 1], table[4 * r + 2]); }
  double mean = 0.0, sd = 1.0;
  if (normalize) {
    mean = s / N;
    double m2 = 0.0;
    for (int r = 0; r < world; ++r) {
      const double nr = table[4 * r], dc = table[4 * r + 1] - mean;
      if (nr > 0.0) m2 += table[4 * r + 3] + 2.0 * dc * table[4 * r + 2] + nr * dc * dc;
    }
    sd = sqrt(m2 / N);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tail3[0] = N;
    if (normalize) { tail3[1] = mean; tail3[2] = sd; }     // normalize = 0: an earlier call's mean / std stay in place
    // a NaN / Inf among the targets (of any rank) shows in the sums: the scan of check_X_y for y
    double chk = s;
    for (int r = 0; r < world; ++r) chk += table[4 * r + 3];
    if (!(fabs(chk) <= 1.7976931348623157e308) || !(fabs(sd) <= 1.7976931348623157e308)) *flag = 1u;
  }
  if (!normalize) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) yt[i] = (y[i] - mean) / sd;
}

// S (lower triangle and diagonal, mirrored) = K(Z, Z) with GPy's exact diagonal + jitter + beta P;  rhs = beta b.
// S holds the kernel entries on entry (lower triangle is what is read); one thread per lower-triangle entry.
__global__ void __launch_bounds__(256) form_system_kernel(double* __restrict__ S, int m, int64_t lds, double sf2, double jitter,
                                                          double beta, const double* __restrict__ P, int64_t ldp,
                                                          const double* __restrict__ b, double* __restrict__ rhs) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m) rhs[idx] = beta * b[idx];
  if (idx >= (int64_t)m * m) return;
  const int r = (int)(idx / m), c = (int)(idx % m);
  if (c > r) return;
  const double k = r == c ? sf2 + jitter : S[r * lds + c];
  const double v = fma(beta, P[r * ldp + c], k);
  S[r * lds + c] = v;
  if (c < r) S[c * lds + r] = v;
}

cudaError_t launch_target_moments(const double* y, int64_t n, double* part, unsigned int* ticket, double* slot, int sms,
                                  cudaStream_t st) {
  int64_t grid = (n + 8 * TM_THREADS - 1) / (8 * TM_THREADS);
  if (grid > 2 * sms) grid = 2 * sms;
  if (grid < 1) grid = 1;
  target_moments_kernel<<<(unsigned)grid, TM_THREADS, 0, st>>>(y, n, part, ticket, slot); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_target_standardize(const double* table, int world, const double* y, int64_t n, double* yt,
                                      double* tail3, int normalize, unsigned int* flag, int sms, cudaStream_t st) {
  int64_t grid = normalize ? (n + 4 * 256 - 1) / (4 * 256) : 1;
  if (grid > 4 * sms) grid = 4 * sms;
  if (grid < 1) grid = 1;
  target_standardize_kernel<<<(unsigned)grid, 256, 0, st>>>(table, world, y, n, yt, tail3, normalize, flag); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_form_system(double* S, int m, int64_t lds, double sf2, double jitter, double beta, const double* P,
                               int64_t ldp, const double* b, double* rhs, cudaStream_t st) {
  form_system_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256, 0, st>>>(S, m, lds, sf2, jitter, beta, P, ldp, b, rhs);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Workspace layout (offsets in doubles, every region 16-byte aligned)
// ---------------------------------------------------------------------------------------------------------------
static inline size_t even(size_t x) { return x + (x & 1); }

size_t fixed_layout(int64_t n, int d, int m, int64_t chunk_rows, int world, int sms, int64_t* off, int stats_mode) {
  const size_t pack = ((size_t)padded_dim(d) + (size_t)((m + MT - 1) / MT) * pack_tile_doubles(padded_dim(d)));
  const int64_t rows = chunk_rows < n ? chunk_rows : n;
  size_t scratch = gemm_tn_workspace_bytes(rows, m, m, 1, sms) / 8;
  const size_t grad = (size_t)sms * padded_dim(d) * padded_dim(d);
  const size_t eig = eigh_workspace_doubles(d);
  const size_t mom = (size_t)4 * sms + 2;
  if (grad > scratch) scratch = grad;
  if (eig > scratch) scratch = eig;
  if (mom > scratch) scratch = mom;
  if (stats_mode == 1) {
    // + 1 KB: the digit planes are placed on a 1 KB boundary inside the scratch region (32-byte sector stores and
    // 16 KB bulk copies; measured with a 16-byte aligned base: 1.33 instead of 0.94 ms for the digit pass)
    const size_t i8 = (inducing_stats_i8_workspace_bytes(rows, m, sms) + 1024) / 8;
    if (i8 > scratch) scratch = i8;
  }
  size_t o = 0;
  auto take = [&](int id, size_t count) { off[id] = (int64_t)o; o += even(count); };
  take(FS_PACK_K, pack);
  take(FS_PACK_G, pack);
  take(FS_YT, (size_t)n);
  take(FS_STATS, (size_t)m * m + m + 1);
  take(FS_TABLE, (size_t)4 * world);
  take(FS_S, (size_t)m * (m + (m & 1)));
  take(FS_L, (size_t)m * m);
  take(FS_RHS, (size_t)m);
  take(FS_ALPHA, (size_t)m);
  take(FS_SCRATCH, scratch);
  take(FS_TAIL, 4);
  take(FS_RESULT, (size_t)d + 2 * (size_t)d * d + 4);
  return o;
}

}  // namespace edrgp
