// Fused FP64 kernels of the sparse-GP EDR hot path for B200 (sm_100a).
//
//   pack_inducing_kernel : Z, l, coef -> tiled inducing pack (one TMA bulk copy per 32-point tile)
//   kuf_kernel<DP>       : Kfu tile = sf2 exp(-r^2/2)  (+ b += Kfu^T y)              [unit K1 / K2b]
//   grad_gram_kernel<DP> : Kfu tile -> W = Kfu o coef -> G = W Z/l^2 - rowsum(W) x/l^2 -> C += G^T G
//                          without ever materialising Kfu, W or (optionally) G       [units K1+K4+K5]
//
// Data layout and roofline are described in DESIGN.md.  Both contractions run on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind).  The inducing tiles are
// staged into shared memory by TMA bulk copies (cp.async.bulk -> SASS UBLKCP) issued by a producer
// warp and handed to the 8 compute warps through an mbarrier full/empty ring; the X row tile is
// double-buffered the same way.  Each compute warp owns 16 data rows: the distance-GEMM
// accumulators (16 x 32 per inducing tile) are turned into W in registers and fed straight back as
// the A operand of the second GEMM by permuting its k index, so nothing but X and (optionally) G
// touches HBM.
#include <cstdlib>
#include "common.cuh"

namespace edrgp {

// -------------------------------------------------------------------------------------------------
// pack layout:  [ il2[DP] | tile 0 | tile 1 | ... ]   tile = MT x S (z/l^2) | MT (hz) | MT (coef)
// -------------------------------------------------------------------------------------------------
__global__ void pack_inducing_kernel(const double* __restrict__ Z, const double* __restrict__ ell,
                                     const double* __restrict__ coef, double coef_scale,
                                     const double* __restrict__ dev_scale, int m, int d,
                                     int dp, double* __restrict__ pack) {
  if (dev_scale != nullptr) coef_scale *= dev_scale[0];     // e.g. std(y), still on the device (no read-back)
  const int S = row_stride(dp);
  const int tile_doubles = pack_tile_doubles(dp);
  const int mtiles = (m + MT - 1) / MT;
  // header: 1 / l^2, zero in the padding
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < dp; q += gridDim.x * blockDim.x) {
    double v = 0.0;
    if (q < d) { double l = ell[q]; v = 1.0 / (l * l); }
    pack[q] = v;
  }
  double* tiles = pack + dp;
  // one warp per inducing point
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < mtiles * MT; j += nwarps) {
    double* tile = tiles + (size_t)(j / MT) * tile_doubles;
    const int jr = j % MT;
    double zn = 0.0;
    for (int q = lane; q < S; q += 32) {
      double v = 0.0;
      if (j < m && q < d) {
        double l = ell[q];
        double z = Z[(size_t)j * d + q];
        double zs = z / l;          // GPy: X2 / lengthscale, then sum(square(.))
        zn = fma(zs, zs, zn);
        v = z / (l * l);
      }
      tile[jr * S + q] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zn += __shfl_xor_sync(0xffffffffu, zn, o);
    if (lane == 0) {
      tile[MT * S + jr] = -0.5 * zn;
      double c = 0.0;
      if (j < m) c = (coef ? coef[j] : 1.0) * coef_scale;
      tile[MT * S + MT + jr] = c;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// shared-memory carve-up (dynamic)
// -------------------------------------------------------------------------------------------------
template <int DP, int XS, int NS>
struct Smem {
  static constexpr int S = row_stride(DP);
  static constexpr int TILE = pack_tile_doubles(DP);
  static constexpr size_t x_off = 0;
  static constexpr size_t z_off = x_off + (size_t)XS * BM * S * 8;
  static constexpr size_t il2_off = z_off + (size_t)NS * TILE * 8;
  static constexpr size_t hx_off = il2_off + (size_t)DP * 8;
  static constexpr size_t tab_off = hx_off + (size_t)BM * 8;             // exp table, replicated per half-warp lane
  static constexpr size_t bar_off = tab_off + (size_t)EXP_TABLE_DOUBLES * 8;
  static constexpr size_t bytes = bar_off + (size_t)(2 * XS + 2 * NS) * 8;
};

struct PipeParams {
  const double* X;
  int64_t n;
  int d;
  const double* pack;
  int mtiles;
  int64_t ntiles;
  double* G;        // may be null
  double* Cpart;    // [gridDim.x][DP*DP], may be null
  // kuf only
  double sf2;
  double* Kfu;
  int64_t ldk;
  int m;
  const double* y;
  double* b;
  double* mu;       // kuf only: mu_i = sum_j K_ij coef_j (posterior mean when coef = alpha)
  const double* Kin;   // cached-Kfu gradient kernel: stored entries (n, ldk), read instead of recomputed
  int64_t ldx;         // leading dimension of X (>= d): a feature block of a wider matrix can be streamed
  int64_t ldg;         // leading dimension of G
  int mul;             // kuf: multiply the entries already in Kfu by this block's factor (feature-chunked d > 128)
  int linear;          // kuf: store the plain contraction x . z / l^2 (projection X V^T) instead of the kernel entry
  unsigned int* flag;  // kuf: set to 1 when a row's scaled norm is not finite (NaN / Inf input: sklearn's check_X_y)
};

// Producer warp: streams the X row tiles and the inducing tiles of every row tile of this CTA.
template <int DP, int XS, int NS>
__device__ __forceinline__ void producer_loop(const PipeParams& p, double* xbuf, double* zbuf, uint64_t* xfull,
                                              uint64_t* xempty, uint64_t* zfull, uint64_t* zempty, int lane) {
  using L = Smem<DP, XS, NS>;
  const double* tiles = p.pack + DP;
  const uint32_t row_bytes = (uint32_t)p.d * 8u;
  int xs = 0, xph = 0, zs = 0, zph = 0;
  auto issue_x = [&](int64_t tile) {
    const int64_t row0 = tile * BM;
    const int rows = (int)min((int64_t)BM, p.n - row0);
    mbar_wait(&xempty[xs], xph ^ 1);
    if (lane == 0) mbar_arrive_expect_tx(&xfull[xs], (uint32_t)rows * row_bytes);
    __syncwarp();
    double* dst = xbuf + (size_t)xs * BM * L::S;
    for (int r = lane; r < rows; r += 32)
      bulk_g2s(dst + (size_t)r * L::S, p.X + (row0 + r) * p.ldx, row_bytes, &xfull[xs]);
    if (++xs == XS) { xs = 0; xph ^= 1; }
  };
  int64_t tile = blockIdx.x;
  if (tile < p.ntiles) issue_x(tile);
  for (; tile < p.ntiles; tile += gridDim.x) {
    for (int mt = 0; mt < p.mtiles; ++mt) {
      mbar_wait(&zempty[zs], zph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&zfull[zs], (uint32_t)L::TILE * 8u);
        bulk_g2s(zbuf + (size_t)zs * L::TILE, tiles + (size_t)mt * L::TILE, (uint32_t)L::TILE * 8u, &zfull[zs]);
      }
      __syncwarp();
      if (++zs == NS) { zs = 0; zph ^= 1; }
      // prefetch the next row tile as early as the ring allows
      if (XS > 1 && mt == 0 && tile + gridDim.x < p.ntiles) issue_x(tile + gridDim.x);
    }
    if (XS == 1 && tile + gridDim.x < p.ntiles) issue_x(tile + gridDim.x);
  }
}

// Producer of the cached-Kfu gradient kernel: streams the pack tiles; the X row tile is only needed
// by the epilogue, so it is requested a few stages into the tile, once the previous tile released it.
// (The stored Kfu entries are read by the compute warps themselves, as register fragments straight
// from global memory: 256-byte row slices are too small for the bulk-copy engine -- one UBLKCP per
// row slice measured 0.6 TB/s.)
template <int DP, int NS>
__device__ __forceinline__ void producer_loop_cached(const PipeParams& p, double* xbuf, double* zbuf, uint64_t* xfull,
                                                     uint64_t* xempty, uint64_t* zfull, uint64_t* zempty, int lane) {
  using L = Smem<DP, 1, NS>;
  const double* tiles = p.pack + DP;
  const uint32_t row_bytes = (uint32_t)p.d * 8u;
  int xph = 0, zs = 0, zph = 0;
  const int x_at = p.mtiles > 2 ? 2 : p.mtiles - 1;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * BM;
    const int rows = (int)min((int64_t)BM, p.n - row0);
    for (int mt = 0; mt < p.mtiles; ++mt) {
      mbar_wait(&zempty[zs], zph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&zfull[zs], (uint32_t)L::TILE * 8u);
        bulk_g2s(zbuf + (size_t)zs * L::TILE, tiles + (size_t)mt * L::TILE, (uint32_t)L::TILE * 8u, &zfull[zs]);
      }
      __syncwarp();
      if (++zs == NS) { zs = 0; zph ^= 1; }
      if (mt == x_at) {
        mbar_wait(&xempty[0], xph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&xfull[0], (uint32_t)rows * row_bytes);
        __syncwarp();
        for (int r = lane; r < rows; r += 32)
          bulk_g2s(xbuf + (size_t)r * L::S, p.X + (row0 + r) * p.ldx, row_bytes, &xfull[0]);
        xph ^= 1;
      }
    }
  }
}

// hx[r] = -0.5 * sum_q (x_rq / l_q)^2 for the 16 rows of this warp (2 lanes per row)
template <int DP>
__device__ __forceinline__ void row_half_norms(const double* xw, const double* il2s, double* hxw, int lane,
                                               unsigned int* flag = nullptr) {
  constexpr int S = row_stride(DP);
  const int r = lane >> 1, h = lane & 1;
  const double* xr = xw + r * S + h * (DP / 2);
  const double* il = il2s + h * (DP / 2);
  double s = 0.0;
#pragma unroll 8
  for (int q = 0; q < DP / 2; ++q) { double x = xr[q]; s = fma(x * x, il[q], s); }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (h == 0) {
    hxw[r] = -0.5 * s;
    // the input scan comes for free here: a NaN or Inf anywhere in the row makes its norm non-finite (rows
    // past the end of the matrix are zero-filled)
    if (flag != nullptr && !(s <= 1.7976931348623157e308)) *flag = 1u;
  }
}

// Distance GEMM for one inducing tile: s[mb][nb][e] = hx + hz + sum_q x_q z_q / l_q^2  (= -r^2/2).
// k is walked in the order {0,1,8,9},{2,3,10,11},... of every 16-feature group so that both
// fragment loads are shared-memory bank-conflict free with row stride == 2 (mod 16).
template <int DP>
__device__ __forceinline__ void dist_gemm(double (&s)[2][MT / 8][2], const double* xr0, const double* xr1,
                                          const double* zt, double hx0, double hx1, int g, int t) {
  constexpr int S = row_stride(DP);
  const double* hz = zt + MT * S;
#pragma unroll
  for (int nb = 0; nb < MT / 8; ++nb) {
    const double2 h = *reinterpret_cast<const double2*>(hz + 8 * nb + 2 * t);
    s[0][nb][0] = hx0 + h.x; s[0][nb][1] = hx0 + h.y;
    s[1][nb][0] = hx1 + h.x; s[1][nb][1] = hx1 + h.y;
  }
  const int cperm = (t & 1) + 8 * (t >> 1);
  const double* zg = zt + g * S + cperm;
  const double* x0 = xr0 + cperm;
  const double* x1 = xr1 + cperm;
#pragma unroll
  for (int kg = 0; kg < DP / 16; ++kg) {
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int col = kg * 16 + 2 * sl;
      const double a0 = x0[col], a1 = x1[col];
#pragma unroll
      for (int nb = 0; nb < MT / 8; ++nb) {
        const double b = zg[nb * 8 * S + col];
        dmma(s[0][nb][0], s[0][nb][1], a0, b);
        dmma(s[1][nb][0], s[1][nb][1], a1, b);
      }
    }
  }
}

// The same contraction in pieces (accumulator start, then one 16-feature group at a time), so that a
// caller can interleave it with independent work at source level.
__device__ __forceinline__ void dist_gemm_init(double (&s)[2][MT / 8][2], const double* hz, double hx0, double hx1, int t) {
#pragma unroll
  for (int nb = 0; nb < MT / 8; ++nb) {
    const double2 h = *reinterpret_cast<const double2*>(hz + 8 * nb + 2 * t);
    s[0][nb][0] = hx0 + h.x; s[0][nb][1] = hx0 + h.y;
    s[1][nb][0] = hx1 + h.x; s[1][nb][1] = hx1 + h.y;
  }
}
template <int DP>
__device__ __forceinline__ void dist_gemm_group(double (&s)[2][MT / 8][2], const double* x0, const double* x1,
                                                const double* zg, int kg) {
  constexpr int S = row_stride(DP);
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const int col = kg * 16 + 2 * sl;
    const double a0 = x0[col], a1 = x1[col];
#pragma unroll
    for (int nb = 0; nb < MT / 8; ++nb) {
      const double b = zg[nb * 8 * S + col];
      dmma(s[0][nb][0], s[0][nb][1], a0, b);
      dmma(s[1][nb][0], s[1][nb][1], a1, b);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// K1+K4+K5 fused
// -------------------------------------------------------------------------------------------------
template <int DP, bool FUSE_GRAM, int XS, int NS, bool FROM_K = false>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1) grad_gram_kernel(const PipeParams p) {
  using L = Smem<DP, XS, NS>;
  static_assert(!FROM_K || XS == 1, "the cached-Kfu variant keeps one X buffer");
  constexpr int S = L::S;
  constexpr int NB = DP / 8;                         // 8-wide feature blocks
  constexpr int CBLK = FUSE_GRAM ? (NB * NB + WARPS - 1) / WARPS : 1;   // Gram blocks per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* xbuf = reinterpret_cast<double*>(smem_raw + L::x_off);
  double* zbuf = reinterpret_cast<double*>(smem_raw + L::z_off);
  double* il2s = reinterpret_cast<double*>(smem_raw + L::il2_off);
  double* hxs = reinterpret_cast<double*>(smem_raw + L::hx_off);
  double* etab_base = reinterpret_cast<double*>(smem_raw + L::tab_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::bar_off);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + XS;
  uint64_t* zfull = bars + 2 * XS;
  uint64_t* zempty = bars + 2 * XS + NS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* etab = etab_base + (lane & 15);       // this lane's copy of the exp table: its own bank pair
  if (tid == 0) {
    for (int i = 0; i < XS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], WARPS); }
    for (int i = 0; i < NS; ++i) { mbar_init(&zfull[i], 1); mbar_init(&zempty[i], WARPS); }
    mbar_fence_init();
  }
  for (int i = tid; i < DP; i += blockDim.x) il2s[i] = p.pack[i];
  exp_table_init(etab_base, tid, blockDim.x);
  // padding columns of the X buffers must be finite zeros (bulk copies fill only d columns)
  for (int i = tid; i < XS * BM * S; i += blockDim.x) xbuf[i] = 0.0;
  fence_proxy_async();
  __syncthreads();

  if (warp == WARPS) {
    if (FROM_K) producer_loop_cached<DP, NS>(p, xbuf, zbuf, xfull, xempty, zfull, zempty, lane);
    else producer_loop<DP, XS, NS>(p, xbuf, zbuf, xfull, xempty, zfull, zempty, lane);
    return;
  }

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * ROWS_PER_WARP;
  double cacc[CBLK][2];
#pragma unroll
  for (int i = 0; i < CBLK; ++i) { cacc[i][0] = 0.0; cacc[i][1] = 0.0; }

  int xs = 0, xph = 0, zs = 0, zph = 0;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * BM;
    double* xt = xbuf + (size_t)xs * BM * S;
    double* xw = xt + r0 * S;
    double hx0 = 0.0, hx1 = 0.0;
    const double* xr0 = xw + g * S;
    const double* xr1 = xw + (g + 8) * S;
    if (!FROM_K) {
      mbar_wait(&xfull[xs], xph);
      row_half_norms<DP>(xw, il2s, hxs + r0, lane);
      __syncwarp();
      hx0 = hxs[r0 + g]; hx1 = hxs[r0 + g + 8];
    }

    double acc[2][NB][2];
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) { acc[0][qb][0] = acc[0][qb][1] = acc[1][qb][0] = acc[1][qb][1] = 0.0; }
    double rs0 = 0.0, rs1 = 0.0;

    // cached variant: this lane's 16 stored Kfu entries of one inducing tile (rows g, g + 8 of the
    // warp; columns 8 nb + {2t, 2t + 1}), fetched one tile ahead straight from global memory
    double2 kf[2][MT / 8];
    const double* krow0 = nullptr;
    const double* krow1 = nullptr;
    auto load_kf = [&](int mt) {
#pragma unroll
      for (int nb = 0; nb < MT / 8; ++nb) {
        const int col = mt * MT + 8 * nb + 2 * t;
        double2 a, b;
        a.x = a.y = b.x = b.y = 0.0;
        if (col < p.m) {       // ldk is even: the pair is in bounds; a column == m carries coefficient 0
          a = __ldg(reinterpret_cast<const double2*>(krow0 + col));
          b = __ldg(reinterpret_cast<const double2*>(krow1 + col));
        }
        kf[0][nb] = a; kf[1][nb] = b;
      }
    };
    if (FROM_K) {
      const int64_t ra = min(row0 + r0 + g, p.n - 1), rb = min(row0 + r0 + g + 8, p.n - 1);   // masked later
      krow0 = p.Kin + ra * p.ldk;
      krow1 = p.Kin + rb * p.ldk;
      load_kf(0);
    }

    for (int mt = 0; mt < p.mtiles; ++mt) {
      const double* zt = zbuf + (size_t)zs * L::TILE;
      mbar_wait(&zfull[zs], zph);
      double s[2][MT / 8][2];
      const double* cf = zt + MT * S + MT;
      if (FROM_K) {
        // W = Kfu * coef from the stored entries; an entry equal to the kernel variance is
        // exp(0): a pair whose clipped r^2 is zero, which GPy's _inv_dist drops
#pragma unroll
        for (int nb = 0; nb < MT / 8; ++nb) {
          const double2 c = *reinterpret_cast<const double2*>(cf + 8 * nb + 2 * t);
          const double2 k0 = kf[0][nb], k1 = kf[1][nb];
          const double w00 = k0.x != p.sf2 ? k0.x * c.x : 0.0, w01 = k0.y != p.sf2 ? k0.y * c.y : 0.0;
          const double w10 = k1.x != p.sf2 ? k1.x * c.x : 0.0, w11 = k1.y != p.sf2 ? k1.y * c.y : 0.0;
          s[0][nb][0] = w00; s[0][nb][1] = w01; s[1][nb][0] = w10; s[1][nb][1] = w11;
          rs0 += w00 + w01; rs1 += w10 + w11;
        }
        if (mt + 1 < p.mtiles) load_kf(mt + 1);       // in flight during the contraction below
      } else {
        dist_gemm<DP>(s, xr0, xr1, zt, hx0, hx1, g, t);
        // W = exp(-r^2/2) * coef, zero where the clipped r^2 is zero (GPy's _inv_dist)
#pragma unroll
        for (int nb = 0; nb < MT / 8; ++nb) {
          const double2 c = *reinterpret_cast<const double2*>(cf + 8 * nb + 2 * t);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const double e0 = s[mb][nb][0], e1 = s[mb][nb][1];
            const double w0 = e0 < 0.0 ? exp_neg(e0, etab) * c.x : 0.0;
            const double w1 = e1 < 0.0 ? exp_neg(e1, etab) * c.y : 0.0;
            s[mb][nb][0] = w0; s[mb][nb][1] = w1;
            if (mb == 0) rs0 += w0 + w1; else rs1 += w0 + w1;
          }
        }
      }
      // acc += W * (Z / l^2): the accumulator columns {2t, 2t+1} of block nb become k-slices whose
      // B rows are the inducing points 8 nb + 2t + sl
#pragma unroll
      for (int nb = 0; nb < MT / 8; ++nb) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const double a0 = s[0][nb][sl], a1 = s[1][nb][sl];
          const double* zr = zt + (8 * nb + 2 * t + sl) * S + g;
#pragma unroll
          for (int qb = 0; qb < NB; ++qb) {
            const double b = zr[8 * qb];
            dmma(acc[0][qb][0], acc[0][qb][1], a0, b);
            dmma(acc[1][qb][0], acc[1][qb][1], a1, b);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&zempty[zs]);
      if (++zs == NS) { zs = 0; zph ^= 1; }
    }

    // G = acc - rowsum(W) * x / l^2, staged over this warp's own X rows
    if (FROM_K) mbar_wait(&xfull[xs], xph);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    const bool v0 = row0 + r0 + g < p.n, v1 = row0 + r0 + g + 8 < p.n;
    double* gr0 = xw + g * S;
    double* gr1 = xw + (g + 8) * S;
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) {
      const int c = 8 * qb + 2 * t;
      const double2 il = *reinterpret_cast<const double2*>(il2s + c);
      const double2 x0 = *reinterpret_cast<const double2*>(gr0 + c);
      const double2 x1 = *reinterpret_cast<const double2*>(gr1 + c);
      double2 o0, o1;
      o0.x = v0 ? fma(-rs0 * x0.x, il.x, acc[0][qb][0]) : 0.0;
      o0.y = v0 ? fma(-rs0 * x0.y, il.y, acc[0][qb][1]) : 0.0;
      o1.x = v1 ? fma(-rs1 * x1.x, il.x, acc[1][qb][0]) : 0.0;
      o1.y = v1 ? fma(-rs1 * x1.y, il.y, acc[1][qb][1]) : 0.0;
      *reinterpret_cast<double2*>(gr0 + c) = o0;
      *reinterpret_cast<double2*>(gr1 + c) = o1;
    }
    __syncwarp();
    if (p.G != nullptr) {
      const int half = p.d >> 1;     // d is even (host guarantees)
      for (int i = lane; i < ROWS_PER_WARP * half; i += 32) {
        const int r = i / half, c2 = i - r * half;
        const int64_t row = row0 + r0 + r;
        if (row < p.n)
          *reinterpret_cast<double2*>(p.G + row * p.ldg + 2 * c2) = *reinterpret_cast<const double2*>(xw + r * S + 2 * c2);
      }
    }
    if (FUSE_GRAM) {
      named_bar_sync(1, WARPS * 32);      // all 128 G rows staged
#pragma unroll 4
      for (int ks = 0; ks < BM / 4; ++ks) {
        const double* gk = xt + (8 * (ks >> 1) + (ks & 1) + 2 * t) * S + g;
#pragma unroll
        for (int i = 0; i < CBLK; ++i) {
          const int id = warp + WARPS * i;
          if (id < NB * NB) {
            const int bi = id / NB, bj = id - bi * NB;
            dmma(cacc[i][0], cacc[i][1], gk[8 * bi], gk[8 * bj]);
          }
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&xempty[xs]);
    if (++xs == XS) { xs = 0; xph ^= 1; }
  }

  if (FUSE_GRAM && p.Cpart != nullptr) {
    double* cp = p.Cpart + (size_t)blockIdx.x * DP * DP;
#pragma unroll
    for (int i = 0; i < CBLK; ++i) {
      const int id = warp + WARPS * i;
      if (id < NB * NB) {
        const int bi = id / NB, bj = id - bi * NB;
        double2 v; v.x = cacc[i][0]; v.y = cacc[i][1];
        *reinterpret_cast<double2*>(cp + (8 * bi + g) * DP + 8 * bj + 2 * t) = v;
      }
    }
  }
}

// C[q][q'] = sum over CTAs of Cpart, cropped from DP to d
__global__ void reduce_gram_kernel(const double* __restrict__ Cpart, int nparts, int dp, int d,
                                   double* __restrict__ C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * d) return;
  const int q = idx / d, r = idx - q * d;
  double s = 0.0;
  for (int i = 0; i < nparts; ++i) s += Cpart[(size_t)i * dp * dp + q * dp + r];
  C[idx] = s;
}

// -------------------------------------------------------------------------------------------------
// K1 (+ b = Kuf y): cross-covariance tiles written to HBM
// -------------------------------------------------------------------------------------------------
// What keeps this kernel at ~62 % of the DMMA peak was looked for in round 2 and NOT found (profiles/r02_kuf_study.txt):
// the data-dependent exp-table lookups were made bank-conflict free (the 40 M "bank conflicts" ncu still counts are
// the TMA writes of the inducing tiles, not LDS wavefronts: per-instruction excess wavefronts are 3 % of the total),
// the FP64 instruction count of the exponential went from 13 to 10, a variant with sixteen warps of 8 rows (four
// warps per scheduler, 92 registers) ran at 60.6 % against 62.2 %, and a register-only microbenchmark shows that two
// warps per scheduler saturate the pipe even with ONE accumulator chain each (tools/dmma_chain.cu), so neither
// occupancy nor accumulator dependencies are the limiter.  DMMA + FP64 pipe time add up to 74 % of the cycles.
// Also tried at the end of round 2 and not kept (same time within 1 %): 128-bit fragment loads (lane t owning the
// features 4t .. 4t+3 of a 16-feature group: half the LDS instructions), and the exponentials' dependent chains woven
// into the DMMA stream one step per three DMMAs (ptxas -O3 re-clusters them; with -O1 the SASS is woven and the time is
// the same), a start-up skew of 400 cycles between the two warps of a scheduler (in case they ran their DMMA and
// exponential bursts in phase).  What did help: not forming the row sums when nobody asks for them (three FP64 instructions per pair).
// SIMPLE: entries are stored, nothing is multiplied in and no Kfu^T y is accumulated -- the statistics
// pass.  Its tile loop is software pipelined: the exp + store of inducing tile t is interleaved, at
// source level and branch-free, with the distance contraction of tile t + 1, so that the DMMA pipe
// is fed while the FP64 exponentials of the previous tile retire (both warps of a sub-partition
// otherwise reach their exp phase together behind the shared tile barrier: 57 % pipe use, ncu r01).
// MU (pipelined variant only): also form the row sums mu = K coef (posterior mean).  Without it the epilogue drops
// three FP64 instructions per entry pair, which run on the pipe the DMMAs need (1.449 -> 1.412 ms per 524 288 rows).
template <int DP, int XS, int NS, bool SIMPLE, bool MU = true>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1) kuf_kernel(const PipeParams p) {
  using L = Smem<DP, XS, NS>;
  constexpr int S = L::S;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* xbuf = reinterpret_cast<double*>(smem_raw + L::x_off);
  double* zbuf = reinterpret_cast<double*>(smem_raw + L::z_off);
  double* il2s = reinterpret_cast<double*>(smem_raw + L::il2_off);
  double* hxs = reinterpret_cast<double*>(smem_raw + L::hx_off);
  double* etab_base = reinterpret_cast<double*>(smem_raw + L::tab_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::bar_off);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + XS;
  uint64_t* zfull = bars + 2 * XS;
  uint64_t* zempty = bars + 2 * XS + NS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* etab = etab_base + (lane & 15);       // this lane's copy of the exp table: its own bank pair
  if (tid == 0) {
    for (int i = 0; i < XS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], WARPS); }
    for (int i = 0; i < NS; ++i) { mbar_init(&zfull[i], 1); mbar_init(&zempty[i], WARPS); }
    mbar_fence_init();
  }
  for (int i = tid; i < DP; i += blockDim.x) il2s[i] = p.pack[i];
  exp_table_init(etab_base, tid, blockDim.x, p.sf2);   // sf2 folded into the table
  for (int i = tid; i < XS * BM * S; i += blockDim.x) xbuf[i] = 0.0;
  fence_proxy_async();
  __syncthreads();

  if (warp == WARPS) {
    producer_loop<DP, XS, NS>(p, xbuf, zbuf, xfull, xempty, zfull, zempty, lane);
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * ROWS_PER_WARP;
  int xs = 0, xph = 0, zs = 0, zph = 0;
  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * BM;
    double* xw = xbuf + (size_t)xs * BM * S + r0 * S;
    mbar_wait(&xfull[xs], xph);
    row_half_norms<DP>(xw, il2s, hxs + r0, lane, p.flag);
    __syncwarp();
    const double hx0 = p.linear ? 0.0 : hxs[r0 + g], hx1 = p.linear ? 0.0 : hxs[r0 + g + 8];
    const int64_t ra = row0 + r0 + g, rb = ra + 8;
    const bool v0 = ra < p.n, v1 = rb < p.n;
    double y0 = 0.0, y1 = 0.0;
    if (p.y != nullptr) { if (v0) y0 = p.y[ra]; if (v1) y1 = p.y[rb]; }
    double mu0 = 0.0, mu1 = 0.0;
    if (SIMPLE) {
      const int cperm = (t & 1) + 8 * (t >> 1);
      const double* x0 = xw + g * S + cperm;
      const double* x1 = xw + (g + 8) * S + cperm;
      double* k0p = p.Kfu + ra * p.ldk + 2 * t;
      double* k1p = p.Kfu + rb * p.ldk + 2 * t;
      // finish one inducing tile: K = sf2 exp(min(s, 0)), row sums, predicated 16-byte stores
      auto finish = [&](double (&sc)[2][MT / 8][2], int nb, int mt, const double* cf) {
        const int j = mt * MT + 8 * nb + 2 * t;
        double2 o0, o1;
        if (p.linear == 2) {       // tuning aid (EDRGP_KUF_DEBUG=1): raw exponents, no exp -- how fast is the loop without it?
          o0.x = sc[0][nb][0]; o0.y = sc[0][nb][1]; o1.x = sc[1][nb][0]; o1.y = sc[1][nb][1];
        } else {
          o0.x = exp_clip_scaled(sc[0][nb][0], etab); o0.y = exp_clip_scaled(sc[0][nb][1], etab);
          o1.x = exp_clip_scaled(sc[1][nb][0], etab); o1.y = exp_clip_scaled(sc[1][nb][1], etab);
        }
        if (MU) {
          const double2 c = *reinterpret_cast<const double2*>(cf + 8 * nb + 2 * t);
          mu0 += fma(o0.x, c.x, o0.y * c.y); mu1 += fma(o1.x, c.x, o1.y * c.y);
        }
        // ldk is even, so a pair starting at an even j < m is in bounds (a column == m is padding)
        if (v0 && j < p.m) *reinterpret_cast<double2*>(k0p + mt * MT + 8 * nb) = o0;
        if (v1 && j < p.m) *reinterpret_cast<double2*>(k1p + mt * MT + 8 * nb) = o1;
      };
      double sa[2][MT / 8][2], sb[2][MT / 8][2];
      const double* zt = zbuf + (size_t)zs * L::TILE;
      mbar_wait(&zfull[zs], zph);
      dist_gemm_init(sa, zt + MT * S, hx0, hx1, t);
#pragma unroll
      for (int kg = 0; kg < DP / 16; ++kg) dist_gemm_group<DP>(sa, x0, x1, zt + g * S + cperm, kg);
      for (int mt = 0; mt + 1 < p.mtiles; ++mt) {
        const double* cf = zt + MT * S + MT;
        const int zs_cur = zs;
        if (++zs == NS) { zs = 0; zph ^= 1; }
        const double* zn = zbuf + (size_t)zs * L::TILE;
        mbar_wait(&zfull[zs], zph);
        dist_gemm_init(sb, zn + MT * S, hx0, hx1, t);
        // DP / 16 contraction groups of the next tile against MT / 8 = 4 finishing steps of this one
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int kg = q * (DP / 16) / 4; kg < (q + 1) * (DP / 16) / 4; ++kg)
            dist_gemm_group<DP>(sb, x0, x1, zn + g * S + cperm, kg);
          finish(sa, q, mt, cf);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&zempty[zs_cur]);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < MT / 8; ++nb) { sa[mb][nb][0] = sb[mb][nb][0]; sa[mb][nb][1] = sb[mb][nb][1]; }
        zt = zn;
      }
      {
        const double* cf = zt + MT * S + MT;
#pragma unroll
        for (int q = 0; q < 4; ++q) finish(sa, q, p.mtiles - 1, cf);
        __syncwarp();
        if (lane == 0) mbar_arrive(&zempty[zs]);
        if (++zs == NS) { zs = 0; zph ^= 1; }
      }
    }
    for (int mt = 0; !SIMPLE && mt < p.mtiles; ++mt) {
      const double* zt = zbuf + (size_t)zs * L::TILE;
      mbar_wait(&zfull[zs], zph);
      double s[2][MT / 8][2];
      dist_gemm<DP>(s, xw + g * S, xw + (g + 8) * S, zt, hx0, hx1, g, t);
      const double* cf = zt + MT * S + MT;
#pragma unroll
      for (int nb = 0; nb < MT / 8; ++nb) {
        const int j = mt * MT + 8 * nb + 2 * t;
        const double2 c = *reinterpret_cast<const double2*>(cf + 8 * nb + 2 * t);
        // entries are stored without the pack coefficient; only the row sums mu carry it
        double k00, k01, k10, k11;
        if (p.linear) {
          // projection mode: the accumulators started from hz only (hx is zeroed above); take hz back out
          const double2 h = *reinterpret_cast<const double2*>(zt + MT * S + 8 * nb + 2 * t);
          k00 = s[0][nb][0] - h.x; k01 = s[0][nb][1] - h.y; k10 = s[1][nb][0] - h.x; k11 = s[1][nb][1] - h.y;
        } else {
          k00 = exp_clip_scaled(s[0][nb][0], etab); k01 = exp_clip_scaled(s[0][nb][1], etab);
          k10 = exp_clip_scaled(s[1][nb][0], etab); k11 = exp_clip_scaled(s[1][nb][1], etab);
        }
        if (p.mul && j < p.m) {
          // feature-chunked evaluation: exp(-r^2/2) factorises over blocks of features
          if (v0) { const double2 o = *reinterpret_cast<const double2*>(p.Kfu + ra * p.ldk + j); k00 *= o.x; k01 *= o.y; }
          if (v1) { const double2 o = *reinterpret_cast<const double2*>(p.Kfu + rb * p.ldk + j); k10 *= o.x; k11 *= o.y; }
        }
        mu0 += fma(k00, c.x, k01 * c.y); mu1 += fma(k10, c.x, k11 * c.y);
        if (p.Kfu != nullptr) {
          // ldk even (host guarantees) and j even -> 16-byte stores
          if (j + 1 < p.m) {
            if (v0) { double2 o; o.x = k00; o.y = k01; *reinterpret_cast<double2*>(p.Kfu + ra * p.ldk + j) = o; }
            if (v1) { double2 o; o.x = k10; o.y = k11; *reinterpret_cast<double2*>(p.Kfu + rb * p.ldk + j) = o; }
          } else if (j < p.m) {
            if (v0) p.Kfu[ra * p.ldk + j] = k00;
            if (v1) p.Kfu[rb * p.ldk + j] = k10;
          }
        }
        if (p.b != nullptr) {
          // column sums of K o y over the warp's 16 rows: reduce over g, then one atomic per column
          double b0 = k00 * y0 + k10 * y1, b1 = k01 * y0 + k11 * y1;
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            b0 += __shfl_xor_sync(0xffffffffu, b0, o);
            b1 += __shfl_xor_sync(0xffffffffu, b1, o);
          }
          if (g == 0 && j < p.m) { atomicAdd(p.b + j, b0); if (j + 1 < p.m) atomicAdd(p.b + j + 1, b1); }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&zempty[zs]);
      if (++zs == NS) { zs = 0; zph ^= 1; }
    }
    if (p.mu != nullptr) {
      mu0 += __shfl_xor_sync(0xffffffffu, mu0, 1); mu0 += __shfl_xor_sync(0xffffffffu, mu0, 2);
      mu1 += __shfl_xor_sync(0xffffffffu, mu1, 1); mu1 += __shfl_xor_sync(0xffffffffu, mu1, 2);
      if (t == 0) { if (v0) p.mu[ra] = mu0; if (v1) p.mu[rb] = mu1; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&xempty[xs]);
    if (++xs == XS) { xs = 0; xph ^= 1; }
  }
}

}  // namespace edrgp

// =================================================================================================
// host-side launchers (called from capi.cu)
// =================================================================================================
#include "launch.h"

namespace edrgp {

template <int DP, bool FUSE, int XS, int NS>
static cudaError_t launch_grad_gram_t(const PipeParams& p, int grid, cudaStream_t st) {
  using L = Smem<DP, XS, NS>;
  auto kern = grad_gram_kernel<DP, FUSE, XS, NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, (WARPS + 1) * 32, L::bytes, st>>>(p); count_launch();
  return cudaGetLastError();
}

template <int DP, int NS>
static cudaError_t launch_grad_gram_cached_t(const PipeParams& p, int grid, cudaStream_t st) {
  using L = Smem<DP, 1, NS>;
  auto kern = grad_gram_kernel<DP, true, 1, NS, true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, (WARPS + 1) * 32, L::bytes, st>>>(p); count_launch();
  return cudaGetLastError();
}

template <int DP, int XS, int NS, bool SIMPLE, bool MU = true>
static cudaError_t launch_kuf_s(const PipeParams& p, int grid, cudaStream_t st) {
  using L = Smem<DP, XS, NS>;
  auto kern = kuf_kernel<DP, XS, NS, SIMPLE, MU>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, (WARPS + 1) * 32, L::bytes, st>>>(p); count_launch();
  return cudaGetLastError();
}

template <int DP, int XS, int NS>
static cudaError_t launch_kuf_t(const PipeParams& p, int grid, cudaStream_t st) {
  // the pipelined variant needs DP / 16 divisible by 4 (64, 128) to split the next tile's contraction
  const bool simple = p.Kfu != nullptr && !p.mul && !p.linear && p.b == nullptr && (DP % 64 == 0);
  static const int dbg = [] { const char* e = getenv("EDRGP_KUF_DEBUG"); return e ? atoi(e) : 0; }();
  if (simple && dbg == 1) { PipeParams q = p; q.linear = 2; return launch_kuf_s<DP, XS, NS, (DP % 64 == 0), false>(q, grid, st); }
  if (simple && p.mu == nullptr) return launch_kuf_s<DP, XS, NS, (DP % 64 == 0), false>(p, grid, st);
  return simple ? launch_kuf_s<DP, XS, NS, (DP % 64 == 0)>(p, grid, st) : launch_kuf_s<DP, XS, NS, false>(p, grid, st);
}

cudaError_t launch_pack(const double* Z, const double* ell, const double* coef, double coef_scale, int m, int d,
                        double* pack, cudaStream_t st, const double* dev_scale) {
  const int dp = padded_dim(d);
  const int mtiles = (m + MT - 1) / MT;
  const int warps_needed = mtiles * MT;
  const int blocks = (warps_needed * 32 + 255) / 256;
  pack_inducing_kernel<<<blocks, 256, 0, st>>>(Z, ell, coef, coef_scale, dev_scale, m, d, dp, pack); count_launch();
  return cudaGetLastError();
}

bool grad_gram_fused(int d) { return padded_dim(d) <= 64; }

cudaError_t launch_grad_gram(const double* X, int64_t n, int d, const double* pack, int m, double* G, double* C,
                             double* Cpart, int sms, cudaStream_t st) {
  PipeParams p{};
  p.X = X; p.n = n; p.d = d; p.pack = pack; p.mtiles = (m + MT - 1) / MT;
  p.ntiles = (n + BM - 1) / BM; p.G = G; p.Cpart = C ? Cpart : nullptr; p.m = m;
  p.ldx = d; p.ldg = d;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  const int dp = padded_dim(d);
  cudaError_t e;
  switch (dp) {
    case 16: e = launch_grad_gram_t<16, true, 2, 4>(p, grid, st); break;
    case 32: e = launch_grad_gram_t<32, true, 2, 4>(p, grid, st); break;
    case 48: e = launch_grad_gram_t<48, true, 2, 4>(p, grid, st); break;
    case 64: e = launch_grad_gram_t<64, true, 2, 4>(p, grid, st); break;
    case 80: e = launch_grad_gram_t<80, false, 1, 3>(p, grid, st); break;
    case 96: e = launch_grad_gram_t<96, false, 1, 3>(p, grid, st); break;
    case 112: e = launch_grad_gram_t<112, false, 1, 2>(p, grid, st); break;
    case 128: e = launch_grad_gram_t<128, false, 1, 2>(p, grid, st); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  if (C != nullptr && dp <= 64) {
    reduce_gram_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(Cpart, grid, dp, d, C); count_launch();
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_grad_gram_cached(const double* X, int64_t ldx, int64_t n, int d, const double* Kin, int64_t ldk,
                                    double sf2, const double* pack, int m, double* G, int64_t ldg, double* C,
                                    double* Cpart, int sms, cudaStream_t st) {
  PipeParams p{};
  p.X = X; p.n = n; p.d = d; p.pack = pack; p.mtiles = (m + MT - 1) / MT;
  p.ntiles = (n + BM - 1) / BM; p.G = G; p.Cpart = C ? Cpart : nullptr; p.m = m;
  p.Kin = Kin; p.ldk = ldk; p.sf2 = sf2; p.ldx = ldx; p.ldg = ldg;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  const int dp = padded_dim(d);
  cudaError_t e;
  switch (dp) {
    case 16: e = launch_grad_gram_cached_t<16, 4>(p, grid, st); break;
    case 32: e = launch_grad_gram_cached_t<32, 4>(p, grid, st); break;
    case 48: e = launch_grad_gram_cached_t<48, 4>(p, grid, st); break;
    case 64: e = launch_grad_gram_cached_t<64, 4>(p, grid, st); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  if (C != nullptr) {
    reduce_gram_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(Cpart, grid, dp, d, C); count_launch();
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_kuf(const double* X, int64_t ldx, int64_t n, int d, const double* pack, int m, double sf2,
                       double* Kfu, int64_t ldk, int mul, const double* y, double* b, double* mu, int sms,
                       cudaStream_t st, int linear, unsigned int* flag) {
  PipeParams p{};
  p.ldx = ldx; p.ldg = d; p.mul = mul; p.linear = linear; p.flag = flag;
  p.X = X; p.n = n; p.d = d; p.pack = pack; p.mtiles = (m + MT - 1) / MT;
  p.ntiles = (n + BM - 1) / BM; p.sf2 = sf2; p.Kfu = Kfu; p.ldk = ldk; p.m = m; p.y = y; p.b = b; p.mu = mu;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  switch (padded_dim(d)) {
    case 16: return launch_kuf_t<16, 2, 4>(p, grid, st);
    case 32: return launch_kuf_t<32, 2, 4>(p, grid, st);
    case 48: return launch_kuf_t<48, 2, 4>(p, grid, st);
    case 64: return launch_kuf_t<64, 2, 4>(p, grid, st);
    case 80: return launch_kuf_t<80, 1, 3>(p, grid, st);
    case 96: return launch_kuf_t<96, 1, 3>(p, grid, st);
    case 112: return launch_kuf_t<112, 1, 2>(p, grid, st);
    case 128: return launch_kuf_t<128, 1, 2>(p, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace edrgp
