// Streaming (HBM-bound) row kernels around the hot path: column moments for the StandardScaler /
// target normaliser, standardisation, and the projection X V^T.
//
// Reference call sites: StandardScaler().fit_transform in
// EffectiveDimensionalityReduction._preprocessing_fit (edrgp/edr.py:161-162), GPy Standardize on y
// (edrgp/gp_model/regression.py:157 with normalizer=True), and EDR.transform
// (edrgp/edr.py:261-289; in-loop use edrgp/base.py:462).
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int CM_THREADS = 256;

// part[blk][0][q] = sum_i w_i (x_iq - shift_q),  part[blk][1][q] = sum_i w_i (x_iq - shift_q)^2
// over the rows of this CTA; deterministic two-stage reduction (no atomics).
__global__ void __launch_bounds__(CM_THREADS) col_moments_kernel(const double* __restrict__ X, int64_t n, int d, int64_t ldx,
                                                                 const double* __restrict__ shift,
                                                                 const double* __restrict__ weight,
                                                                 double* __restrict__ part) {
  extern __shared__ double sh[];     // [CM_THREADS / cols_lanes][2][d] folded below
  // thread layout: lanes walk the columns (coalesced), thread rows walk the rows
  const int lanes = d < 32 ? d : 32;                 // threads per row slice
  const int rows_per_pass = CM_THREADS / lanes;
  const int tr = threadIdx.x / lanes, tc = threadIdx.x % lanes;
  const bool active = tr < rows_per_pass;
  const int ncol = (d + lanes - 1) / lanes;          // columns per thread (<= 16 for d <= 512)
  double s1[16], s2[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
  if (active) {
    for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + tr; r < n; r += (int64_t)gridDim.x * rows_per_pass) {
      const double* xr = X + r * ldx;
      const double w = weight ? weight[r] : 1.0;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int q = tc + c * lanes;
        if (c < ncol && q < d) {
          const double v = xr[q] - (shift ? shift[q] : 0.0);
          const double wv = w * v;
          s1[c] += wv;
          s2[c] = fma(wv, v, s2[c]);
        }
      }
    }
  }
  // fold the row-threads through shared memory, column by column
  double* red = sh;                                   // [CM_THREADS]
  for (int c = 0; c < ncol; ++c) {
    for (int k = 0; k < 2; ++k) {
      double v = 0.0;
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) if (cc == c) v = k == 0 ? s1[cc] : s2[cc];
      red[threadIdx.x] = active ? v : 0.0;
      __syncthreads();
      // pairwise tree over the row-threads (a serial sum by one thread was a chain of CM_THREADS / lanes
      // dependent FP64 additions: 10 us of a 54 us call for a single column)
      for (int cnt = rows_per_pass; cnt > 1;) {
        const int half = (cnt + 1) / 2;
        if (active && tr < cnt / 2) red[tr * lanes + tc] += red[(tr + half) * lanes + tc];
        __syncthreads();
        cnt = half;
      }
      if (tr == 0 && active) {
        const int q = tc + c * lanes;
        if (q < d) part[((size_t)blockIdx.x * 2 + k) * d + q] = red[tc];
      }
      __syncthreads();
    }
  }
}

// one warp per output (2 d of them): lanes stride over the CTA partials with two accumulators, then a
// shuffle tree; fixed order, deterministic
__global__ void __launch_bounds__(256) col_moments_reduce_kernel(const double* __restrict__ part, int nblk, int d,
                                                                 double* __restrict__ out, int64_t ostride, int accumulate) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= 2 * d) return;
  const int k = i / d, q = i - k * d;
  double s0 = 0.0, s1 = 0.0;
  int b = lane;
  for (; b + 32 < nblk; b += 64) {
    s0 += part[((size_t)b * 2 + k) * d + q];
    s1 += part[((size_t)(b + 32) * 2 + k) * d + q];
  }
  if (b < nblk) s0 += part[((size_t)b * 2 + k) * d + q];
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) { double* o = out + k * ostride + q; *o = accumulate ? *o + s : s; }
}

__global__ void standardize_kernel(const double* __restrict__ X, int64_t total, int d, const double* __restrict__ mean,
                                   const double* __restrict__ scale, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % d);
    out[i] = (X[i] - mean[q]) / scale[q];
  }
}

// out[i][c] = sum_q X[i][q] V[c][q]; one warp per row, k <= 16 components per launch
template <int KMAX>
__global__ void __launch_bounds__(256) project_kernel(const double* __restrict__ X, int64_t n, int d,
                                                      const double* __restrict__ V, int k, int64_t ldv,
                                                      double* __restrict__ out, int64_t ldo) {
  extern __shared__ double vs[];    // [k][d]
  for (int i = threadIdx.x; i < k * d; i += blockDim.x) vs[i] = V[(int64_t)(i / d) * ldv + (i % d)];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n; r += nwarps) {
    double acc[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) acc[c] = 0.0;
    const double* xr = X + r * d;
    for (int q = lane; q < d; q += 32) {
      const double x = xr[q];
#pragma unroll
      for (int c = 0; c < KMAX; ++c) if (c < k) acc[c] = fma(x, vs[c * d + q], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
      if (c < k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < KMAX; ++c) if (c < k) out[r * ldo + c] = acc[c];
    }
  }
}

// count[0] += number of non-finite entries (NaN or +-Inf): sklearn's check_X_y / check_array scan
// (edrgp/gp_model/base.py:87,105) as one HBM-speed pass on the device.
__global__ void __launch_bounds__(256) count_nonfinite_kernel(const double* __restrict__ X, int64_t total,
                                                              unsigned int* __restrict__ count) {
  unsigned int bad = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) {
    const int64_t pairs = total >> 1;
    const double2* X2 = reinterpret_cast<const double2*>(X);
    for (int64_t j = i; j < pairs; j += stride) {
      const double2 v = X2[j];
      bad += !isfinite(v.x);
      bad += !isfinite(v.y);
    }
    if ((total & 1) && i == 0) bad += !isfinite(X[total - 1]);
  } else {
    for (int64_t j = i; j < total; j += stride) bad += !isfinite(X[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

cudaError_t launch_count_nonfinite(const double* X, int64_t total, unsigned int* count, int sms, cudaStream_t st) {
  int64_t blocks = (total / 2 + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  count_nonfinite_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, total, count); count_launch();
  return cudaGetLastError();
}

size_t col_moments_workspace_bytes(int d, int sms) { return (size_t)sms * 4 * 2 * d * sizeof(double); }

cudaError_t launch_col_moments(const double* X, int64_t n, int d, const double* shift, const double* weight,
                               double* out, int accumulate, double* workspace, int sms, cudaStream_t st) {
  // a thread keeps 16 column sums: inputs wider than 512 columns are walked in blocks of 512 (out = [2][d])
  for (int c0 = 0; c0 < d; c0 += 512) {
    const int dc = d - c0 < 512 ? d - c0 : 512;
    int grid = sms * 4;
    const int lanes = dc < 32 ? dc : 32;
    const int rows_per_pass = CM_THREADS / lanes;
    const int64_t need = (n + rows_per_pass - 1) / rows_per_pass;
    if (grid > need) grid = (int)need;
    col_moments_kernel<<<grid, CM_THREADS, CM_THREADS * sizeof(double), st>>>(X + c0, n, dc, d, shift ? shift + c0 : nullptr,
                                                                                weight, workspace); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    col_moments_reduce_kernel<<<(2 * dc * 32 + 255) / 256, 256, 0, st>>>(workspace, grid, dc, out + c0, d, accumulate); count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_standardize(const double* X, int64_t n, int d, const double* mean, const double* scale, double* out,
                               int sms, cudaStream_t st) {
  const int64_t total = n * d;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  standardize_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, total, d, mean, scale, out); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_project(const double* X, int64_t n, int d, const double* V, int k, double* out, int sms,
                           cudaStream_t st) {
  int64_t blocks = (n + 7) / 8;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  for (int c0 = 0; c0 < k; c0 += 16) {
    const int kc = k - c0 < 16 ? k - c0 : 16;
    const size_t smem = (size_t)kc * d * sizeof(double);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(project_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    if (kc <= 4) {
      project_kernel<4><<<(unsigned)blocks, 256, smem, st>>>(X, n, d, V + (int64_t)c0 * d, kc, d, out + c0, k);
    } else {
      project_kernel<16><<<(unsigned)blocks, 256, smem, st>>>(X, n, d, V + (int64_t)c0 * d, kc, d, out + c0, k);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace edrgp
