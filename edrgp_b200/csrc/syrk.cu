// C = A^T B for tall row-major A (n x ka), B (n x kb), n >> ka, kb, on the FP64 tensor pipe
// (DMMA.8x8x4), with a symmetric mode C = A^T A that computes only the upper-triangle tiles.
//
// Used for the inducing statistics P = Kuf Kfu, b = Kuf y, y^T y (unit K2: symmetric mode with the
// optional y vector), for the gradient Gram matrix G^T G when d > 64 (unit K5), and for T^T X in the
// hyper-parameter gradient of the VFE bound.  Output tiles are 128 x 128; the n rows are split over
// CTAs (split-K) so that tiles x splits fills the chip, each CTA writing its partial tile to a
// workspace that a second kernel sums in a fixed order (deterministic, no atomics).  Row chunks of
// 16 are staged through a 4-deep cp.async ring; the shared-memory row stride 132 == 4 (mod 16)
// makes both DMMA fragment loads bank-conflict free.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int TB = 128;        // output tile edge
constexpr int KC = 16;         // rows per pipeline stage
constexpr int SST = 4;         // stages
constexpr int SS = TB + 4;     // smem row stride (doubles)
constexpr int STAGE_DOUBLES = 2 * KC * SS + KC;   // A rows, B rows, y slice

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct GemmTnParams {
  const double* A;
  const double* B;
  int64_t n;
  int ka, kb;
  int64_t lda, ldb;
  int sym;         // B == A, only tiles ti <= tj
  int nta, ntb;    // tiles per edge
  int ntiles;
  int ksplit;
  int64_t rows_per_split;   // multiple of KC
  double* part;    // [ksplit][ka * kb]
  const double* y; // optional (sym only): bpart[split][0..ka) = A^T y, bpart[split][ka] = y^T y
  double* bpart;   // [ksplit][ka + 1]
};

__global__ void __launch_bounds__(256, 1) gemm_tn_kernel(const GemmTnParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);   // [SST][STAGE_DOUBLES]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  int tile = blockIdx.x % p.ntiles, split = blockIdx.x / p.ntiles;
  int ti, tj;
  if (p.sym) {
    ti = 0;
    int rem = tile;
    while (rem >= p.ntb - ti) { rem -= p.ntb - ti; ++ti; }
    tj = ti + rem;
  } else {
    ti = tile / p.ntb;
    tj = tile - ti * p.ntb;
  }
  const bool diag = p.sym && ti == tj;
  const bool with_y = diag && p.y != nullptr;
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = min(p.n, r_begin + p.rows_per_split);
  const int nchunks = r_end > r_begin ? (int)((r_end - r_begin + KC - 1) / KC) : 0;

  const int wm = warp >> 1, wn = warp & 1;          // warp tile: rows 32*wm.., cols 64*wn..
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double bacc = 0.0;      // tid < 128: column ti*TB + tid of A^T y;  tid == 128: y^T y

  // loader: each stage = 2 operands x KC rows x 128 doubles = 2 x 16 x 64 16-byte pieces (+ 8 for y)
  auto load_stage = [&](int chunk, int stage) {
    const int64_t row_base = r_begin + (int64_t)chunk * KC;
    double* base = sm + (size_t)stage * STAGE_DOUBLES;
#pragma unroll
    for (int it = 0; it < (2 * KC * (TB / 2)) / 256; ++it) {
      const int idx = tid + it * 256;
      const int op = idx / (KC * (TB / 2));
      if (op == 1 && diag) continue;
      const int rr = (idx / (TB / 2)) % KC;
      const int c2 = idx % (TB / 2);
      const int col = (op == 0 ? ti : tj) * TB + 2 * c2;
      const int kk = op == 0 ? p.ka : p.kb;
      const int64_t row = row_base + rr;
      // ka, kb and the leading dimensions are even: a 16-byte piece never straddles the edge
      const bool ok = row < r_end && col < kk;
      const double* mat = op == 0 ? p.A : p.B;
      const double* src = ok ? mat + row * (op == 0 ? p.lda : p.ldb) + col : mat;
      cp_async16(base + (size_t)op * KC * SS + rr * SS + 2 * c2, src, ok ? 16 : 0);
    }
    if (with_y && tid < KC / 2) {
      const int64_t row = row_base + 2 * tid;
      const int bytes = row + 1 < r_end ? 16 : (row < r_end ? 8 : 0);
      cp_async16(base + 2 * KC * SS + 2 * tid, bytes ? p.y + row : p.y, bytes);
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
    const double* As = sm + (size_t)(c % SST) * STAGE_DOUBLES;
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
    if (with_y) {
      const double* ys = As + 2 * KC * SS;
      if (tid < TB) {
#pragma unroll
        for (int r = 0; r < KC; ++r) bacc = fma(As[r * SS + tid], ys[r], bacc);
      } else if (tid == TB && ti == 0) {
#pragma unroll
        for (int r = 0; r < KC; ++r) bacc = fma(ys[r], ys[r], bacc);
      }
    }
  }
  cp_async_wait<0>();

  double* out = p.part + (size_t)split * p.ka * p.kb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ti * TB + 32 * wm + 8 * i + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cidx = tj * TB + 64 * wn + 8 * j + 2 * t;
      if (r < p.ka && cidx < p.kb) {
        out[(size_t)r * p.kb + cidx] = acc[i][j][0];
        if (cidx + 1 < p.kb) out[(size_t)r * p.kb + cidx + 1] = acc[i][j][1];
      }
    }
  }
  if (with_y) {
    double* bo = p.bpart + (size_t)split * (p.ka + 1);
    if (tid < TB) {
      if (ti * TB + tid < p.ka) bo[ti * TB + tid] = bacc;
    } else if (tid == TB && ti == 0) {
      bo[p.ka] = bacc;
    }
  }
}

// C (+)= sum over splits of the partials; in symmetric mode the lower triangle mirrors the upper.
__global__ void gemm_tn_reduce_kernel(const double* __restrict__ part, int ksplit, int ka, int kb, int sym,
                                      int accumulate, double* __restrict__ C, int64_t ldc,
                                      const double* __restrict__ bpart, double* __restrict__ bout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)ka * kb;
  if (idx < total) {
    const int r = (int)(idx / kb), c = (int)(idx % kb);
    // symmetric: element (r, c) lives in tile (r/TB, c/TB); only tiles with ti <= tj were computed
    const bool direct = !sym || (r / TB) <= (c / TB);
    const size_t src = direct ? (size_t)r * kb + c : (size_t)c * kb + r;
    double s = 0.0;
    for (int i = 0; i < ksplit; ++i) s += part[(size_t)i * total + src];
    double* dst = C + (int64_t)r * ldc + c;
    *dst = accumulate ? *dst + s : s;
  } else if (bout != nullptr && idx < total + ka + 1) {
    const int j = (int)(idx - total);
    double s = 0.0;
    for (int i = 0; i < ksplit; ++i) s += bpart[(size_t)i * (ka + 1) + j];
    bout[j] = accumulate ? bout[j] + s : s;
  }
}

static void gemm_tn_plan(int64_t n, int ka, int kb, int sym, int sms, GemmTnParams* p) {
  p->nta = (ka + TB - 1) / TB;
  p->ntb = (kb + TB - 1) / TB;
  p->ntiles = sym ? p->ntb * (p->ntb + 1) / 2 : p->nta * p->ntb;
  int ks = sms / p->ntiles;
  if (ks < 1) ks = 1;
  const int64_t chunks = (n + KC - 1) / KC;
  if (ks > chunks) ks = (int)chunks;
  const int64_t cps = (chunks + ks - 1) / ks;
  p->rows_per_split = cps * KC;
  p->ksplit = (int)((n + p->rows_per_split - 1) / p->rows_per_split);
}

size_t gemm_tn_workspace_bytes(int64_t n, int ka, int kb, int sym, int sms) {
  GemmTnParams p{};
  gemm_tn_plan(n, ka, kb, sym, sms, &p);
  return ((size_t)p.ksplit * ka * kb + (size_t)p.ksplit * (ka + 1)) * sizeof(double);
}

cudaError_t launch_gemm_tn(const double* A, int64_t lda, int ka, const double* B, int64_t ldb, int kb, int64_t n,
                           int sym, const double* y, double* C, int64_t ldc, double* bout, int accumulate,
                           double* workspace, int sms, cudaStream_t st) {
  GemmTnParams p{};
  p.A = A; p.B = sym ? A : B; p.n = n; p.ka = ka; p.kb = sym ? ka : kb; p.lda = lda; p.ldb = sym ? lda : ldb;
  p.sym = sym; p.y = sym ? y : nullptr;
  gemm_tn_plan(n, p.ka, p.kb, sym, sms, &p);
  p.part = workspace;
  p.bpart = workspace + (size_t)p.ksplit * p.ka * p.kb;
  const size_t smem = (size_t)SST * STAGE_DOUBLES * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gemm_tn_kernel<<<p.ntiles * p.ksplit, 256, smem, st>>>(p); count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)p.ka * p.kb + (p.y ? p.ka + 1 : 0);
  gemm_tn_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(workspace, p.ksplit, p.ka, p.kb, sym,
                                                                         accumulate, C, ldc, p.bpart,
                                                                         p.y ? bout : nullptr); count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
