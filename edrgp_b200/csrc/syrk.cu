// C = A^T A for a tall row-major A (n x k, n >> k) on the FP64 tensor pipe (DMMA.8x8x4).
//
// Used for the inducing statistic P = Kuf Kfu (k = m, unit K2) and for the gradient Gram matrix
// G^T G when d > 64 (unit K5).  Output tiles are 128 x 128 (upper triangle only, mirrored by the
// reduce kernel); the n rows are split over CTAs (split-K) so that tiles x splits fills the chip,
// each CTA writing its partial tile to a workspace that a second kernel sums in a fixed order
// (deterministic, no atomics).  Row chunks of 16 are staged through a 4-deep cp.async ring; the
// shared-memory row stride 132 == 4 (mod 16) makes both DMMA fragment loads bank-conflict free.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int TB = 128;        // output tile edge
constexpr int KC = 16;         // rows per pipeline stage
constexpr int SST = 4;         // stages
constexpr int SS = TB + 4;     // smem row stride (doubles)

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SyrkParams {
  const double* A;
  int64_t n;
  int k;
  int64_t lda;
  int nt;          // tiles per edge
  int ntiles;      // upper-triangle tiles
  int ksplit;
  int64_t rows_per_split;   // multiple of KC
  double* part;    // [ksplit][k][k] (only upper tiles written)
};

__global__ void __launch_bounds__(256, 1) syrk_tn_kernel(const SyrkParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);   // [SST][2][KC][SS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // tile index -> (ti <= tj)
  int tile = blockIdx.x % p.ntiles, split = blockIdx.x / p.ntiles;
  int ti = 0, rem = tile;
  while (rem >= p.nt - ti) { rem -= p.nt - ti; ++ti; }
  const int tj = ti + rem;
  const bool diag = ti == tj;
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = min(p.n, r_begin + p.rows_per_split);
  const int nchunks = r_end > r_begin ? (int)((r_end - r_begin + KC - 1) / KC) : 0;

  const int wm = warp >> 1, wn = warp & 1;          // warp tile: rows 32*wm.., cols 64*wn..
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // loader: each stage = 2 operands x KC rows x 128 doubles = 2 x 16 x 64 16-byte pieces
  auto load_stage = [&](int chunk, int stage) {
    const int64_t row_base = r_begin + (int64_t)chunk * KC;
    double* base = sm + (size_t)stage * 2 * KC * SS;
#pragma unroll
    for (int it = 0; it < (2 * KC * (TB / 2)) / 256; ++it) {
      const int idx = tid + it * 256;
      const int op = idx / (KC * (TB / 2));
      if (op == 1 && diag) continue;
      const int rr = (idx / (TB / 2)) % KC;
      const int c2 = idx % (TB / 2);
      const int col = (op == 0 ? ti : tj) * TB + 2 * c2;
      const int64_t row = row_base + rr;
      const bool ok = row < r_end && col < p.k;      // k even: a 16-byte piece never straddles k
      const double* src = ok ? p.A + row * p.lda + col : p.A;
      cp_async16(base + (size_t)op * KC * SS + rr * SS + 2 * c2, src, ok ? 16 : 0);
    }
  };

  for (int s = 0; s < SST - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<SST - 2>();
    __syncthreads();
    if (c + SST - 1 < nchunks) load_stage(c + SST - 1, (c + SST - 1) % SST);
    cp_async_commit();
    const double* As = sm + (size_t)(c % SST) * 2 * KC * SS;
    const double* Bs = diag ? As : As + KC * SS;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      const double* ar = As + (4 * ks + t) * SS + 32 * wm + g;
      const double* br = Bs + (4 * ks + t) * SS + 64 * wn + g;
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ar[8 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = br[8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  double* out = p.part + (size_t)split * p.k * p.k;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ti * TB + 32 * wm + 8 * i + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cidx = tj * TB + 64 * wn + 8 * j + 2 * t;
      if (r < p.k && cidx < p.k) {
        out[(size_t)r * p.k + cidx] = acc[i][j][0];
        if (cidx + 1 < p.k) out[(size_t)r * p.k + cidx + 1] = acc[i][j][1];
      }
    }
  }
}

// C = sum over splits of the upper-tile partials, mirrored to the lower triangle.
__global__ void syrk_reduce_kernel(const double* __restrict__ part, int ksplit, int k, double* __restrict__ C,
                                   int64_t ldc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)k * k) return;
  const int r = (int)(idx / k), c = (int)(idx % k);
  // element (r, c) lives in tile (r/TB, c/TB); only tiles with ti <= tj were computed
  const bool upper = (r / TB) <= (c / TB);
  const size_t src = upper ? (size_t)r * k + c : (size_t)c * k + r;
  double s = 0.0;
  for (int i = 0; i < ksplit; ++i) s += part[(size_t)i * k * k + src];
  C[(int64_t)r * ldc + c] = s;
}

static void syrk_plan(int64_t n, int k, int sms, int* nt, int* ntiles, int* ksplit, int64_t* rows_per_split) {
  *nt = (k + TB - 1) / TB;
  *ntiles = *nt * (*nt + 1) / 2;
  int ks = sms / *ntiles;
  if (ks < 1) ks = 1;
  const int64_t chunks = (n + KC - 1) / KC;
  if (ks > chunks) ks = (int)chunks;
  int64_t cps = (chunks + ks - 1) / ks;
  *rows_per_split = cps * KC;
  *ksplit = (int)((n + *rows_per_split - 1) / *rows_per_split);
}

size_t syrk_workspace_bytes(int64_t n, int k, int sms) {
  int nt, ntiles, ksplit; int64_t rps;
  syrk_plan(n, k, sms, &nt, &ntiles, &ksplit, &rps);
  return (size_t)ksplit * k * k * sizeof(double);
}

cudaError_t launch_syrk(const double* A, int64_t n, int k, int64_t lda, double* C, int64_t ldc, double* workspace,
                        int sms, cudaStream_t st) {
  SyrkParams p{};
  p.A = A; p.n = n; p.k = k; p.lda = lda; p.part = workspace;
  syrk_plan(n, k, sms, &p.nt, &p.ntiles, &p.ksplit, &p.rows_per_split);
  const size_t smem = (size_t)SST * 2 * KC * SS * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(syrk_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  syrk_tn_kernel<<<p.ntiles * p.ksplit, 256, smem, st>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)k * k;
  syrk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(workspace, p.ksplit, k, C, ldc);
  return cudaGetLastError();
}

}  // namespace edrgp
