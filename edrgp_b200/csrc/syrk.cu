// C = A^T B for tall row-major A (n x ka), B (n x kb), n >> ka, kb, on the FP64 tensor pipe
// (DMMA.8x8x4), with a symmetric mode C = A^T A that computes only the upper-triangle tiles.
//
// Used for the inducing statistics P = Kuf Kfu, b = Kuf y, y^T y (unit K2: symmetric mode with the
// optional y vector), for the gradient Gram matrix G^T G when d > 64 (unit K5), and for T^T X in the
// hyper-parameter gradient of the VFE bound.  Output tiles are 128 x 128; the n rows are split over
// CTAs (split-K) so that tiles x splits fills the chip, each CTA writing its partial tile to a
// workspace that a second kernel sums in a fixed order (deterministic, no atomics).  Row chunks of
// 32 are staged through a 3-deep ring by a TMA producer warp (cp.async.bulk, one row slice per
// lane) with mbarrier full/empty hand-off; the shared-memory row stride 132 == 4 (mod 16) makes both
// DMMA fragment loads bank-conflict free.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int TB = 128;        // output tile edge
constexpr int KC = 32;         // rows per pipeline stage (one producer lane per row)
constexpr int SST = 3;         // stages
constexpr int SS = TB + 4;     // smem row stride (doubles)
constexpr int STAGE_DOUBLES = 2 * KC * SS + KC;   // A rows, B rows, y slice
constexpr int GT_WARPS = 8;    // compute warps; warp 8 is the TMA producer
constexpr size_t GT_BAR_OFF = (size_t)SST * STAGE_DOUBLES * sizeof(double);
constexpr size_t GT_SMEM = GT_BAR_OFF + 2 * SST * sizeof(uint64_t);

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

// Operand staging is done by a dedicated producer warp with TMA bulk copies (one 1 KB row slice
// per lane and operand, SASS UBLKCP) completing on an mbarrier full/empty ring: the 8 compute warps
// issue nothing but LDS + DMMA in the main loop.  (With cp.async issued by the compute warps the
// address arithmetic of all warps lines up behind each barrier and idles the DMMA pipe ~20 % of the
// time: tools/dmma_limits.cu, profiles/r01_ncu_top_kernels.txt.)
__global__ void __launch_bounds__((GT_WARPS + 1) * 32, 1) gemm_tn_kernel(const GemmTnParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);   // [SST][STAGE_DOUBLES]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + GT_BAR_OFF);
  uint64_t* empty = full + SST;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

  if (tid == 0) {
    for (int i = 0; i < SST; ++i) { mbar_init(&full[i], 32); mbar_init(&empty[i], GT_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == GT_WARPS) {
    // ------------------------------ producer ------------------------------
    // 16-byte granules: an odd column count is rounded up (the leading dimensions are even, so the
    // extra column is in bounds; it only feeds output rows / columns that are never written)
    const int acols = min(TB, p.ka - ti * TB), bcols = min(TB, p.kb - tj * TB);
    const uint32_t abytes = (uint32_t)((acols + 1) & ~1) * 8u, bbytes = diag ? 0u : (uint32_t)((bcols + 1) & ~1) * 8u;
    double yy = 0.0;
    int stage = 0, ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int64_t row_base = r_begin + (int64_t)c * KC;
      const int valid = (int)min((int64_t)KC, r_end - row_base);
      double* base = sm + (size_t)stage * STAGE_DOUBLES;
      mbar_wait(&empty[stage], ph ^ 1);
      const int64_t row = row_base + lane;
      const bool ok = lane < valid;
      if (with_y) {
        const double yv = ok ? p.y[row] : 0.0;
        base[2 * KC * SS + lane] = yv;
        if (ti == 0) yy = fma(yv, yv, yy);
      }
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)valid * (abytes + bbytes));
      __syncwarp();
      if (ok) {
        bulk_g2s(base + lane * SS, p.A + row * p.lda + (int64_t)ti * TB, abytes, &full[stage]);
        if (!diag) bulk_g2s(base + KC * SS + lane * SS, p.B + row * p.ldb + (int64_t)tj * TB, bbytes, &full[stage]);
      } else {
        // ragged last chunk: rows past the end must read as zeros
        for (int q = 0; q < TB; ++q) { base[lane * SS + q] = 0.0; if (!diag) base[KC * SS + lane * SS + q] = 0.0; }
      }
      if (lane != 0) mbar_arrive(&full[stage]);
      if (++stage == SST) { stage = 0; ph ^= 1; }
    }
    if (with_y && ti == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) yy += __shfl_xor_sync(0xffffffffu, yy, o);
      if (lane == 0) p.bpart[(size_t)split * (p.ka + 1) + p.ka] = yy;
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;          // warp tile: rows 32*wm.., cols 64*wn..
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double bacc = 0.0;      // column ti*TB + (tid & 127) of A^T y over the rows tid >> 7 of each half stage

  int stage = 0, ph = 0;
  for (int c = 0; c < nchunks; ++c) {
    const double* As = sm + (size_t)stage * STAGE_DOUBLES;
    const double* Bs = diag ? As : As + KC * SS;
    mbar_wait(&full[stage], ph);
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
      const double* ys = As + 2 * KC * SS + (tid >> 7) * (KC / 2);
      const double* ac = As + (tid >> 7) * (KC / 2) * SS + (tid & (TB - 1));
#pragma unroll
      for (int r = 0; r < KC / 2; ++r) bacc = fma(ac[r * SS], ys[r], bacc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == SST) { stage = 0; ph ^= 1; }
  }

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
    // fold the two row halves: all loads of the ring are done once every warp passed its last wait
    named_bar_sync(1, GT_WARPS * 32);
    double* ex = sm;          // [128]
    if (tid >= TB) ex[tid - TB] = bacc;
    named_bar_sync(1, GT_WARPS * 32);
    if (tid < TB && ti * TB + tid < p.ka) p.bpart[(size_t)split * (p.ka + 1) + ti * TB + tid] = bacc + ex[tid];
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
  const size_t smem = GT_SMEM;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gemm_tn_kernel<<<p.ntiles * p.ksplit, (GT_WARPS + 1) * 32, smem, st>>>(p); count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)p.ka * p.kb + (p.y ? p.ka + 1 : 0);
  gemm_tn_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(workspace, p.ksplit, p.ka, p.kb, sym,
                                                                         accumulate, C, ldc, p.bpart,
                                                                         p.y ? bout : nullptr); count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
