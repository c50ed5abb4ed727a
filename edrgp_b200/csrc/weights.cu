// T = K o (c_ya * y alpha^T + c_km * K M)  for a stored cross-covariance block K (n x m) and a
// symmetric-or-not m x m matrix M, with row sums and per-row-block column sums of T, on the FP64
// tensor pipe (DMMA.8x8x4).
//
// This is the n-scale contraction of the VFE-bound hyper-parameter gradient: with
// M = dL/dpsi2 and c_ya = beta, c_km = 2 it yields T = Kfu o dL/dKfu, from which
//   dL/d(variance) = sum(T) / variance,   dL/dZ = (T^T X - colsum(T) o Z) / l^2,
//   dL/dl_q = (sum_i rowsum(T)_i x_iq^2 - 2 sum_j z_jq (T^T X)_jq + sum_j colsum(T)_j z_jq^2) / l_q^3
// (GPy Stationary.update_gradients_full / gradients_X applied to dL_dKnm of VarDTC.inference,
// reached from model.optimize at edrgp/gp_model/base.py:69).  With M = woodbury_inv, c_ya = 0,
// c_km = 1 the row sums are k_i^T W k_i of the predictive variance (GPy Posterior._raw_predict,
// edrgp/gp_model/base.py:206).
//
// Tiling: 128 x 128 output tiles, k-loop over m in chunks of 16 through a 4-deep cp.async ring.
// A = K rows are kept [row][k] with stride 20 == 4 (mod 16), B = M rows [k][col] with stride 132
// == 4 (mod 16): both DMMA fragment loads are bank-conflict free.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

namespace {
constexpr int WT = 128;        // tile edge
constexpr int WK = 16;         // k per stage
constexpr int WST = 4;         // stages
constexpr int SA = WK + 4;     // A smem row stride
constexpr int SB = WT + 4;     // B smem row stride
constexpr int WSTAGE = WT * SA + WK * SB;

__device__ __forceinline__ void cpa16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace

struct WeightsParams {
  const double* K; int64_t n; int m; int64_t ldk;
  const double* M; int64_t ldm;
  const double* y; const double* alpha;     // may be null when c_ya == 0
  double c_ya, c_km;
  double* T; int64_t ldt;                   // may be null
  double* rs_part;                          // [ntj][n] or null
  double* cs_part;                          // [nrb][m] or null
  int ntj;
};

__global__ void __launch_bounds__(256, 1) weights_kernel(const WeightsParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tj = blockIdx.x % p.ntj;
  const int64_t rb = blockIdx.x / p.ntj;
  const int64_t row0 = rb * WT;
  const int col0 = tj * WT;
  const int wm = warp >> 1, wn = warp & 1;
  const int nk = (p.m + WK - 1) / WK;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int kc, int stage) {
    double* As = sm + (size_t)stage * WSTAGE;
    double* Bs = As + WT * SA;
    const int k0 = kc * WK;
#pragma unroll
    for (int it = 0; it < 4; ++it) {          // A: 128 rows x 8 pieces
      const int idx = tid + it * 256;
      const int r = idx >> 3, c2 = idx & 7;
      const int64_t row = row0 + r;
      const int k = k0 + 2 * c2;
      const bool ok = row < p.n && k < p.m;    // ldk even: pieces never straddle the edge
      cpa16(As + r * SA + 2 * c2, ok ? p.K + row * p.ldk + k : p.K, ok ? 16 : 0);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {          // B: 16 k-rows x 64 pieces
      const int idx = tid + it * 256;
      const int kr = idx >> 6, c2 = idx & 63;
      const int k = k0 + kr, col = col0 + 2 * c2;
      const bool ok = k < p.m && col < p.m;
      cpa16(Bs + kr * SB + 2 * c2, ok ? p.M + (int64_t)k * p.ldm + col : p.M, ok ? 16 : 0);
    }
  };

  for (int s = 0; s < WST - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cpa_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    cpa_wait<WST - 2>();
    __syncthreads();
    if (kc + WST - 1 < nk) load_stage(kc + WST - 1, (kc + WST - 1) % WST);
    cpa_commit();
    const double* As = sm + (size_t)(kc % WST) * WSTAGE;
    const double* Bs = As + WT * SA;
#pragma unroll
    for (int ks = 0; ks < WK / 4; ++ks) {
      const double* ar = As + (32 * wm + g) * SA + 4 * ks + t;
      const double* br = Bs + (4 * ks + t) * SB + 64 * wn + g;
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ar[8 * i * SA];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = br[8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cpa_wait<0>();
  __syncthreads();       // the ring is free: reuse it for the row/column sum exchange

  double* rs_s = sm;               // [2][128]
  double* cs_s = sm + 2 * WT;      // [4][128]
  double csum[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j][0] = csum[j][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = 32 * wm + 8 * i + g;
    const int64_t row = row0 + rl;
    const bool rok = row < p.n;
    const double yv = (rok && p.y != nullptr) ? p.c_ya * p.y[row] : 0.0;
    double rsum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + 64 * wn + 8 * j + 2 * t;
      double t0 = 0.0, t1 = 0.0;
      if (rok && col < p.m) {
        // m may be odd only in the last column; ldk even keeps the 16-byte load in bounds
        const double2 kv = *reinterpret_cast<const double2*>(p.K + row * p.ldk + col);
        double a0 = 0.0, a1 = 0.0;
        if (p.alpha != nullptr) { a0 = p.alpha[col]; a1 = col + 1 < p.m ? p.alpha[col + 1] : 0.0; }
        t0 = kv.x * fma(p.c_km, acc[i][j][0], yv * a0);
        t1 = col + 1 < p.m ? kv.y * fma(p.c_km, acc[i][j][1], yv * a1) : 0.0;
        if (p.T != nullptr) {
          double2 o; o.x = t0; o.y = t1;
          *reinterpret_cast<double2*>(p.T + row * p.ldt + col) = o;
        }
      }
      rsum += t0 + t1;
      csum[j][0] += t0; csum[j][1] += t1;
    }
    rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
    rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
    if (t == 0) rs_s[wn * WT + rl] = rsum;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = csum[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) cs_s[wm * WT + 64 * wn + 8 * j + 2 * t + e] = v;
    }
  }
  __syncthreads();
  if (tid < WT) {
    if (p.rs_part != nullptr && row0 + tid < p.n)
      p.rs_part[(size_t)tj * p.n + row0 + tid] = rs_s[tid] + rs_s[WT + tid];
  } else if (tid < 2 * WT) {
    const int c = tid - WT;
    if (p.cs_part != nullptr && col0 + c < p.m)
      p.cs_part[(size_t)rb * p.m + col0 + c] = (cs_s[c] + cs_s[WT + c]) + (cs_s[2 * WT + c] + cs_s[3 * WT + c]);
  }
}

// rowsum[i] = sum_tj rs_part[tj][i];  colsum[c] (+)= sum_rb cs_part[rb][c]  (fixed order)
__global__ void weights_reduce_kernel(const double* __restrict__ rs_part, int ntj, int64_t n, double* __restrict__ rowsum,
                                      const double* __restrict__ cs_part, int64_t nrb, int m, double* __restrict__ colsum,
                                      int accumulate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) {
    if (rowsum != nullptr) {
      double s = 0.0;
      for (int k = 0; k < ntj; ++k) s += rs_part[(size_t)k * n + idx];
      rowsum[idx] = s;
    }
  } else if (idx < n + m) {
    if (colsum != nullptr) {
      const int c = (int)(idx - n);
      double s = 0.0;
      for (int64_t r = 0; r < nrb; ++r) s += cs_part[(size_t)r * m + c];
      colsum[c] = accumulate ? colsum[c] + s : s;
    }
  }
}

size_t weights_workspace_bytes(int64_t n, int m) {
  const int ntj = (m + WT - 1) / WT;
  const int64_t nrb = (n + WT - 1) / WT;
  return ((size_t)ntj * n + (size_t)nrb * m) * sizeof(double);
}

cudaError_t launch_weights(const double* K, int64_t n, int m, int64_t ldk, const double* M, int64_t ldm, const double* y,
                           const double* alpha, double c_ya, double c_km, double* T, int64_t ldt, double* rowsum,
                           double* colsum, int accumulate, double* workspace, cudaStream_t st) {
  WeightsParams p{};
  p.K = K; p.n = n; p.m = m; p.ldk = ldk; p.M = M; p.ldm = ldm; p.y = y; p.alpha = alpha;
  p.c_ya = (y && alpha) ? c_ya : 0.0; p.c_km = c_km; p.T = T; p.ldt = ldt;
  p.ntj = (m + WT - 1) / WT;
  const int64_t nrb = (n + WT - 1) / WT;
  p.rs_part = rowsum ? workspace : nullptr;
  p.cs_part = colsum ? workspace + (size_t)p.ntj * n : nullptr;
  const size_t smem = (size_t)WST * WSTAGE * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (nrb * p.ntj > 0x7fffffffLL) return cudaErrorInvalidValue;
  weights_kernel<<<(unsigned)(nrb * p.ntj), 256, smem, st>>>(p); count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (rowsum || colsum) {
    const int64_t total = n + m;
    weights_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.rs_part, p.ntj, n, rowsum, p.cs_part, nrb, m,
                                                                         colsum, accumulate); count_launch();
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace edrgp
