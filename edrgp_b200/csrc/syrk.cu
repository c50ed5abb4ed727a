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
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <vector>
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
  int bn;          // B-side tile width: 128, or 64 in general mode when kb <= 64
  int nta, ntb;    // tiles per edge
  int ntiles;      // general mode: nta * ntb; symmetric mode: off-diagonal tiles ntb (ntb - 1) / 2
  int ksplit;      // row splits of the general / off-diagonal tiles
  int64_t rows_per_split;   // multiple of KC
  int ndiag;       // symmetric mode: diagonal tiles (= ntb), computed as triangles (17/32 of the DMMAs)
  int ksplit_diag;
  int64_t rows_per_split_diag;
  double* part;    // [ksplit][ka * kb]
  const double* y; // optional (sym only): bpart[split][0..ka) = A^T y, bpart[split][ka] = y^T y
  double* bpart;   // [ksplit][ka + 1]
};

// Consumer of a DIAGONAL tile of the symmetric mode: only the upper triangle of the 16 x 16 grid of
// 8 x 8 output blocks is computed.  Warp W owns block rows W and 15 - W (16 - W and W + 1 blocks:
// 17 DMMAs per k-step for every warp instead of 32), so a diagonal tile costs 17/32 of an
// off-diagonal one and gets proportionally fewer row splits (gemm_tn_plan).
template <int W>
__device__ __forceinline__ void diag_consumer(const GemmTnParams& p, const double* sm, uint64_t* full, uint64_t* empty,
                                              int nchunks, int ti, int split, bool with_y) {
  constexpr int NLO = 16 - W, NHI = W + 1, RLO = W, RHI = 15 - W;
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  double lo[NLO][2], hi[NHI][2];
#pragma unroll
  for (int j = 0; j < NLO; ++j) lo[j][0] = lo[j][1] = 0.0;
#pragma unroll
  for (int j = 0; j < NHI; ++j) hi[j][0] = hi[j][1] = 0.0;
  double bacc = 0.0;      // column ti*TB + (tid & 127) of A^T y over the rows tid >> 7 of each half stage
  int stage = 0, ph = 0;
  for (int c = 0; c < nchunks; ++c) {
    const double* As = sm + (size_t)stage * STAGE_DOUBLES;
    mbar_wait(&full[stage], ph);
    // the eight warps run eight different instantiations of this loop: keep each body small, or the
    // instruction cache thrashes (ncu: stall_no_instruction dominated with the loop fully unrolled)
#pragma unroll 2
    for (int ks = 0; ks < KC / 4; ++ks) {
      const double* row = As + (4 * ks + t) * SS + g;
      double b[NLO];
#pragma unroll
      for (int j = 0; j < NLO; ++j) b[j] = row[8 * (RLO + j)];
      const double alo = b[0], ahi = b[RHI - RLO];       // the A operand is the same column slab
#pragma unroll
      for (int j = 0; j < NLO; ++j) dmma(lo[j][0], lo[j][1], alo, b[j]);
#pragma unroll
      for (int j = 0; j < NHI; ++j) dmma(hi[j][0], hi[j][1], ahi, b[RHI - RLO + j]);
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
  const int r_lo = ti * TB + 8 * RLO + g, r_hi = ti * TB + 8 * RHI + g;
#pragma unroll
  for (int j = 0; j < NLO; ++j) {
    const int cidx = ti * TB + 8 * (RLO + j) + 2 * t;
    if (r_lo < p.ka && cidx < p.kb) {
      out[(size_t)r_lo * p.kb + cidx] = lo[j][0];
      if (cidx + 1 < p.kb) out[(size_t)r_lo * p.kb + cidx + 1] = lo[j][1];
    }
  }
#pragma unroll
  for (int j = 0; j < NHI; ++j) {
    const int cidx = ti * TB + 8 * (RHI + j) + 2 * t;
    if (r_hi < p.ka && cidx < p.kb) {
      out[(size_t)r_hi * p.kb + cidx] = hi[j][0];
      if (cidx + 1 < p.kb) out[(size_t)r_hi * p.kb + cidx + 1] = hi[j][1];
    }
  }
  if (with_y) {
    // fold the two row halves: all loads of the ring are done once every warp passed its last wait
    named_bar_sync(1, GT_WARPS * 32);
    double* ex = const_cast<double*>(sm);          // [128]
    if (tid >= TB) ex[tid - TB] = bacc;
    named_bar_sync(1, GT_WARPS * 32);
    if (tid < TB && ti * TB + tid < p.ka) p.bpart[(size_t)split * (p.ka + 1) + ti * TB + tid] = bacc + ex[tid];
  }
}

// Consumer of a full tile: 128 rows of C x BN = 16 NJ columns (BN = 128, or 64 for a narrow B such as
// the (n, d <= 64) data rows of T^T X, which would leave half of a 128-wide tile empty).
template <int NJ>
__device__ __forceinline__ void full_consumer(const GemmTnParams& p, const double* sm, uint64_t* full, uint64_t* empty,
                                              int nchunks, int ti, int tj, int split) {
  constexpr int BN = 16 * NJ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;          // warp tile: rows 32*wm.., cols (BN/2)*wn..
  double acc[4][NJ][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  int stage = 0, ph = 0;
  for (int c = 0; c < nchunks; ++c) {
    const double* As = sm + (size_t)stage * STAGE_DOUBLES;
    const double* Bs = As + KC * SS;
    mbar_wait(&full[stage], ph);
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      const double* ar = As + (4 * ks + t) * SS + 32 * wm + g;
      const double* br = Bs + (4 * ks + t) * SS + (BN / 2) * wn + g;
      double a[4], b[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ar[8 * i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) b[j] = br[8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
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
    for (int j = 0; j < NJ; ++j) {
      const int cidx = tj * BN + (BN / 2) * wn + 8 * j + 2 * t;
      if (r < p.ka && cidx < p.kb) {
        out[(size_t)r * p.kb + cidx] = acc[i][j][0];
        if (cidx + 1 < p.kb) out[(size_t)r * p.kb + cidx + 1] = acc[i][j][1];
      }
    }
  }
}

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
  // CTA -> (tile, row split): the off-diagonal / general tiles come first, then the diagonal ones
  const int noff = p.ntiles * p.ksplit;
  int ti, tj, split;
  int64_t rps;
  bool diag = false;
  if ((int)blockIdx.x < noff) {
    const int tile = blockIdx.x % p.ntiles;
    split = blockIdx.x / p.ntiles;
    rps = p.rows_per_split;
    if (p.sym) {             // pairs ti < tj
      ti = 0;
      int rem = tile;
      while (rem >= p.ntb - 1 - ti) { rem -= p.ntb - 1 - ti; ++ti; }
      tj = ti + 1 + rem;
    } else {
      ti = tile / p.ntb;
      tj = tile - ti * p.ntb;
    }
  } else {
    const int idx = blockIdx.x - noff;
    ti = tj = idx % p.ndiag;
    split = idx / p.ndiag;
    rps = p.rows_per_split_diag;
    diag = true;
  }
  const bool with_y = diag && p.y != nullptr;
  const int64_t r_begin = (int64_t)split * rps;
  const int64_t r_end = min(p.n, r_begin + rps);
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
    const int acols = min(TB, p.ka - ti * TB), bcols = min(p.bn, p.kb - tj * p.bn);
    const uint32_t abytes = (uint32_t)((acols + 1) & ~1) * 8u, bbytes = diag ? 0u : (uint32_t)((bcols + 1) & ~1) * 8u;
    double yy = 0.0;
    int stage = 0, ph = 0;
    // the y slice of a chunk is fetched one chunk ahead so that its latency never sits between the
    // release of a stage and the bulk copies that refill it
    double ynext = (with_y && r_begin + lane < r_end) ? p.y[r_begin + lane] : 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const int64_t row_base = r_begin + (int64_t)c * KC;
      const int valid = (int)min((int64_t)KC, r_end - row_base);
      double* base = sm + (size_t)stage * STAGE_DOUBLES;
      const double yv = ynext;
      if (with_y) {
        const int64_t nrow = row_base + KC + lane;
        ynext = nrow < r_end ? p.y[nrow] : 0.0;
      }
      mbar_wait(&empty[stage], ph ^ 1);
      const int64_t row = row_base + lane;
      const bool ok = lane < valid;
      if (with_y) {
        base[2 * KC * SS + lane] = yv;
        if (ti == 0) yy = fma(yv, yv, yy);
      }
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)valid * (abytes + bbytes));
      __syncwarp();
      if (ok) {
        bulk_g2s(base + lane * SS, p.A + row * p.lda + (int64_t)ti * TB, abytes, &full[stage]);
        if (!diag) bulk_g2s(base + KC * SS + lane * SS, p.B + row * p.ldb + (int64_t)tj * p.bn, bbytes, &full[stage]);
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
  if (diag) {
    switch (warp) {
      case 0: diag_consumer<0>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 1: diag_consumer<1>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 2: diag_consumer<2>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 3: diag_consumer<3>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 4: diag_consumer<4>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 5: diag_consumer<5>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      case 6: diag_consumer<6>(p, sm, full, empty, nchunks, ti, split, with_y); break;
      default: diag_consumer<7>(p, sm, full, empty, nchunks, ti, split, with_y); break;
    }
    return;
  }
  if (p.bn == TB) full_consumer<8>(p, sm, full, empty, nchunks, ti, tj, split);
  else full_consumer<4>(p, sm, full, empty, nchunks, ti, tj, split);
}

// C (+)= sum over splits of the partials; in symmetric mode only the 8 x 8 blocks on or above the
// block diagonal were computed (diagonal tiles are triangles) and the rest is mirrored.
__global__ void gemm_tn_reduce_kernel(const double* __restrict__ part, int ksplit, int ksplit_diag, int ka, int kb,
                                      int sym, int accumulate, double* __restrict__ C, int64_t ldc,
                                      const double* __restrict__ bpart, double* __restrict__ bout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)ka * kb;
  if (idx < total) {
    const int r = (int)(idx / kb), c = (int)(idx % kb);
    const bool direct = !sym || (r >> 3) <= (c >> 3);
    const int rr = direct ? r : c, cc = direct ? c : r;
    const int ns = (sym && rr / TB == cc / TB) ? ksplit_diag : ksplit;
    const size_t src = (size_t)rr * kb + cc;
    double s = 0.0;
    for (int i = 0; i < ns; ++i) s += part[(size_t)i * total + src];
    double* dst = C + (int64_t)r * ldc + c;
    *dst = accumulate ? *dst + s : s;
  } else if (bout != nullptr && idx < total + ka + 1) {
    const int j = (int)(idx - total);
    double s = 0.0;
    for (int i = 0; i < ksplit_diag; ++i) s += bpart[(size_t)i * (ka + 1) + j];
    bout[j] = accumulate ? bout[j] + s : s;
  }
}

static void split_rows(int64_t n, int want, int* ksplit, int64_t* rows_per_split) {
  const int64_t chunks = (n + KC - 1) / KC;
  int ks = want < 1 ? 1 : want;
  if (ks > chunks) ks = (int)chunks;
  const int64_t cps = (chunks + ks - 1) / ks;
  *rows_per_split = cps * KC;
  *ksplit = (int)((n + *rows_per_split - 1) / *rows_per_split);
}

// (ko, kd) for a symmetric reduction with `ntiles` full and `ndiag` triangular tiles; memoised per
// (ntb, sms) -- the search simulates the block scheduler and is not free.
static void choose_splits(int ntiles, int ndiag, int sms, int* ko_out, int* kd_out) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, std::pair<int, int>> memo;
  std::lock_guard<std::mutex> lock(mu);
  auto it = memo.find({ndiag, sms});
  if (it != memo.end()) { *ko_out = it->second.first; *kd_out = it->second.second; return; }
  const double cdiag = 17.0 / 32.0;
  int best_off = ntiles ? 1 : 0, best_diag = 1;
  double best = 1e300;
  const int kmax = ntiles ? 64 : 4 * sms;
  std::vector<double> busy(sms);
  for (int kd = 1; kd <= kmax; ++kd) {
    for (int ko = ntiles ? 1 : 0; ko <= (ntiles ? 64 : 0); ++ko) {
      const int noff = ntiles * ko, ctas = noff + ndiag * kd;
      if (ctas > 8 * sms) break;
      // per-CTA cost: its share of the rows plus a small fixed prologue / epilogue term
      const double coff = ntiles ? 1.0 / ko + 2e-4 : 0.0, cdg = cdiag / kd + 2e-4;
      // the equal-cost full tiles land round-robin; the triangular ones then fill greedily
      for (int q = 0; q < sms; ++q) busy[q] = coff * (noff / sms + (q < noff % sms ? 1 : 0));
      std::make_heap(busy.begin(), busy.end(), std::greater<double>());
      for (int c = 0; c < ndiag * kd; ++c) {
        std::pop_heap(busy.begin(), busy.end(), std::greater<double>());
        busy.back() += cdg;
        std::push_heap(busy.begin(), busy.end(), std::greater<double>());
      }
      const double t = *std::max_element(busy.begin(), busy.end());
      if (t < best) { best = t; best_off = ko; best_diag = kd; }
    }
  }
  memo[{ndiag, sms}] = {best_off, best_diag};
  *ko_out = best_off; *kd_out = best_diag;
}

static void gemm_tn_plan(int64_t n, int ka, int kb, int sym, int sms, GemmTnParams* p) {
  p->bn = (!sym && kb <= TB / 2) ? TB / 2 : TB;
  p->nta = (ka + TB - 1) / TB;
  p->ntb = (kb + p->bn - 1) / p->bn;
  if (!sym) {
    p->ntiles = p->nta * p->ntb;
    p->ndiag = 0; p->ksplit_diag = 0; p->rows_per_split_diag = KC;
    split_rows(n, sms / p->ntiles, &p->ksplit, &p->rows_per_split);
    return;
  }
  // symmetric: ntb (ntb - 1) / 2 full tiles (cost 1 per row) and ntb triangular ones (cost 17/32).
  // Pick the split counts (ko, kd) with the smallest makespan of the hardware's list scheduling
  // (CTAs in launch order -- full tiles first -- each to the first SM that frees up, one per SM).
  p->ntiles = p->ntb * (p->ntb - 1) / 2;
  p->ndiag = p->ntb;
  int best_off, best_diag;
  choose_splits(p->ntiles, p->ndiag, sms, &best_off, &best_diag);
  if (p->ntiles) split_rows(n, best_off, &p->ksplit, &p->rows_per_split);
  else { p->ksplit = 0; p->rows_per_split = KC; }
  split_rows(n, best_diag, &p->ksplit_diag, &p->rows_per_split_diag);
}

size_t gemm_tn_workspace_bytes(int64_t n, int ka, int kb, int sym, int sms) {
  GemmTnParams p{};
  gemm_tn_plan(n, ka, kb, sym, sms, &p);
  const size_t slots = (size_t)(p.ksplit > p.ksplit_diag ? p.ksplit : p.ksplit_diag);
  return (slots * ka * kb + slots * (ka + 1)) * sizeof(double);
}

cudaError_t launch_gemm_tn(const double* A, int64_t lda, int ka, const double* B, int64_t ldb, int kb, int64_t n,
                           int sym, const double* y, double* C, int64_t ldc, double* bout, int accumulate,
                           double* workspace, int sms, cudaStream_t st) {
  GemmTnParams p{};
  p.A = A; p.B = sym ? A : B; p.n = n; p.ka = ka; p.kb = sym ? ka : kb; p.lda = lda; p.ldb = sym ? lda : ldb;
  p.sym = sym; p.y = sym ? y : nullptr;
  gemm_tn_plan(n, p.ka, p.kb, sym, sms, &p);
  p.part = workspace;
  const size_t slots = (size_t)(p.ksplit > p.ksplit_diag ? p.ksplit : p.ksplit_diag);
  p.bpart = workspace + slots * p.ka * p.kb;
  const size_t smem = GT_SMEM;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gemm_tn_kernel<<<p.ntiles * p.ksplit + p.ndiag * p.ksplit_diag, (GT_WARPS + 1) * 32, smem, st>>>(p); count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)p.ka * p.kb + (p.y ? p.ka + 1 : 0);
  gemm_tn_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(workspace, p.ksplit, p.ksplit_diag, p.ka, p.kb, sym,
                                                                         accumulate, C, ldc, p.bpart,
                                                                         p.y ? bout : nullptr); count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
