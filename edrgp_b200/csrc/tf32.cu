// TF32-split ("tf32x3") cross-covariance for B200 (sm_100a): the optional reduced-precision mode of
// unit K1 (BASELINE north star: kernel entries within 1e-4 relative of the FP64 path).
//
//   Kfu[i][j] = sf2 * exp(-0.5 * max(|x_i/l|^2 + |z_j/l|^2 - 2 (x_i/l).(z_j/l), 0))
//
// The distance contraction (x/l).(z/l) runs on the 5th-generation tensor cores:
// tcgen05.mma.kind::tf32 (SASS UTCMMA) with both operands in shared memory and the FP32
// accumulators in tensor memory.  Each FP64 operand is split into two TF32 values,
// v = hi + lo (22 significant bits), and three products are accumulated, lo*hi + hi*lo + hi*hi
// (the dropped lo*lo term is ~2^-22 relative).  The norms are kept outside the contraction in FP64
// and enter the exponent as two FP32 addends.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      producer : streams the packed inducing chunks (128 points: hi | lo images, already in
//                          the 128-byte-swizzled K-major layout the MMA reads) with TMA bulk copies
//   warp 1      MMA      : allocates the 512 TMEM columns; one thread issues the tcgen05.mma sequence
//                          (M = 128 rows, N = 128 inducing points, K = 8 per instruction) and commits
//                          to the mbarriers that free the operand buffers / publish the accumulators
//   warps 2-9   convert  : read the FP64 rows of the X tile (coalesced, one row per warp pass), scale
//                          by 1/l, split into hi / lo and store them swizzled; row norms -> ring
//   warps 10-17 epilogue : tcgen05.ld the accumulators (a thread per row, 32 columns per load),
//                          exponent + ex2.approx + FP32->FP64, 32-byte vector stores of Kfu rows
// Four accumulator stages of 128 columns let the contraction of chunk j + 1 .. j + 3 run under the
// epilogue of chunk j: the kernel is bound by the 8 m bytes per point it has to write to HBM.
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace edrgp {

namespace tf32 {

constexpr int TM = 128;                    // data rows per tile (UMMA M)
constexpr int TN = 128;                    // inducing points per chunk (UMMA N)
constexpr int KBLK = 32;                   // TF32 elements per 128-byte swizzled row
constexpr int BLK_BYTES = TM * 128;        // one operand k-block: 128 rows x 128 B = 16 KB
constexpr int MAX_KB = 2;                  // d <= 64
constexpr int ACC_STAGES = 4;              // 4 x 128 = 512 TMEM columns
constexpr int Z_STAGES = 1;                // inducing chunk buffers (64 KB each at d = 64)
constexpr int X_STAGES = 2;                // converted X tiles (64 KB each at d = 64)
constexpr int XN_RING = 8;                 // row-norm ring (tiles): converters run <= 5 tiles ahead of the epilogue
constexpr int CONV_WARPS = 8;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 32 * (2 + CONV_WARPS + EPI_WARPS);
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  const double* X;
  int64_t ldx;
  int64_t n;
  int d;                 // features (even, <= 64)
  const double* ell;     // lengthscales (d)
  const uint8_t* pack;   // [cz: mpad floats | chunk images]
  int m;
  int nchunks;
  int kb;                // k-blocks of 32 features
  int ksteps;            // MMA k-steps of 8 features
  float l2sf2;           // log2(sf2)
  double sf2;            // the value stored where r^2 clips at 0 (exactly sf2, as in the FP64 kernel)
  double* K;
  int64_t ldk;
  int64_t ntiles;
  int debug;             // tuning aid (EDRGP_TF32_DEBUG): 1 = no global stores, 2 = no X conversion after the first tile
};

__host__ __device__ inline int chunk_bytes(int kb) { return 2 * kb * BLK_BYTES; }
__host__ __device__ inline int mpad(int m) { return (m + TN - 1) / TN * TN; }
__host__ __device__ inline size_t cz_bytes(int m) { return ((size_t)mpad(m) * 4 + 1023) / 1024 * 1024; }

// byte offset of element (row r, feature k within the 32-wide block) in a 128-byte-swizzled K-major block
__host__ __device__ inline uint32_t sw128_offset(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4);
}

__device__ __forceinline__ uint32_t to_tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
// v = hi + lo with hi, lo representable in TF32
__device__ __forceinline__ void split_tf32(double v, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32((float)v);
  lo = to_tf32((float)(v - (double)__uint_as_float(hi)));
}
// The same split for the hot conversion loops WITHOUT FP64 <-> FP32 conversion instructions: those run on
// the transcendental (XU) pipe at 16 lanes per clock and SM, and three of them per element made that pipe
// the limiter of the gradient kernel (ncu: XU saturated, HBM at 43 %).  Veltkamp's multiplication by
// 2^42 + 1 rounds v to 11 significant bits in FP64 (exactly a TF32 value); FP32 bit patterns are then
// assembled from the double's words with integer instructions.  The low part keeps its spare mantissa
// bits (the tensor core reads the TF32 field of the 32-bit container).  |v| < 2^-126 flushes to zero.
__device__ __forceinline__ uint32_t f32_bits_of(double x) {
  const int h = __double2hiint(x);
  const uint32_t l = (uint32_t)__double2loint(x);
  const int a = h & 0x7fffffff;
  uint32_t f = ((uint32_t)(a - 0x38000000) << 3) | (l >> 29);
  f = a < 0x38100000 ? 0u : f;
  return f | ((uint32_t)h & 0x80000000u);
}
// FP32 bits of a double that carries at most 21 significant bits (only its high word matters)
__device__ __forceinline__ uint32_t f32_bits_of_short(double x) {
  const int h = __double2hiint(x);
  const int a = h & 0x7fffffff;
  const uint32_t f = (uint32_t)(a - 0x38000000) << 3;
  return (a < 0x38100000 ? 0u : f) | ((uint32_t)h & 0x80000000u);
}
// v >= 0 (kernel entries): no sign handling for the high part; the low part keeps 20 mantissa bits
__device__ __forceinline__ void split_tf32_bits_nonneg(double v, uint32_t& hi, uint32_t& lo) {
  const double t = __dmul_rn(v, 4398046511105.0);
  const double hd = __dsub_rn(t, __dsub_rn(t, v));
  const int a = __double2hiint(hd);
  hi = a < 0x38100000 ? 0u : (uint32_t)(a - 0x38000000) << 3;
  lo = f32_bits_of_short(__dsub_rn(v, hd));
}
__device__ __forceinline__ void split_tf32_bits(double v, uint32_t& hi, uint32_t& lo) {
  // (explicitly rounded operations: an FMA contraction of t - v would undo the rounding this relies on)
  const double t = __dmul_rn(v, 4398046511105.0);          // 2^42 + 1
  const double hd = __dsub_rn(t, __dsub_rn(t, v));         // v rounded to 11 significant bits
  hi = f32_bits_of(hd);
  lo = f32_bits_of(__dsub_rn(v, hd));
}

// ---- tcgen05 wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address, 16-byte units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused with swizzle): 1
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_v4_f64(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---- inducing pack: cz[j] = -0.5 |z_j/l|^2 log2(e) (FP32), then per 128-point chunk the hi and lo images ----
__global__ void pack_tf32_kernel(const double* __restrict__ Z, const double* __restrict__ ell, int m, int d, int kb,
                                 uint8_t* __restrict__ pack) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int mp = mpad(m);
  if (warp >= mp) return;
  const int j = warp;
  float* cz = reinterpret_cast<float*>(pack);
  uint8_t* chunk = pack + cz_bytes(m) + (size_t)(j / TN) * chunk_bytes(kb);
  const int r = j % TN;
  double zn = 0.0;
  for (int q = lane; q < kb * KBLK; q += 32) {
    double zs = 0.0;
    if (j < m && q < d) zs = Z[(size_t)j * d + q] / ell[q];            // GPy: X2 / lengthscale
    zn = fma(zs, zs, zn);
    uint32_t hi, lo;
    split_tf32(zs, hi, lo);
    const uint32_t off = (uint32_t)(q / KBLK) * BLK_BYTES + sw128_offset(r, q % KBLK);
    *reinterpret_cast<uint32_t*>(chunk + off) = hi;
    *reinterpret_cast<uint32_t*>(chunk + (size_t)kb * BLK_BYTES + off) = lo;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) zn += __shfl_xor_sync(0xffffffffu, zn, o);
  if (lane == 0) cz[j] = (float)(-0.5 * zn * 1.4426950408889634);
}

struct __align__(8) Barriers {
  uint64_t z_full[Z_STAGES], z_empty[Z_STAGES];
  uint64_t x_full[X_STAGES], x_empty[X_STAGES];
  uint64_t acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
};

__global__ void __launch_bounds__(THREADS, 1) kuf_tf32_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (the swizzle pattern is a function of the absolute shared address)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int kb = p.kb;
  uint8_t* sX = smem;                                        // X_STAGES x [hi kb blocks | lo kb blocks]
  uint8_t* sZ = sX + X_STAGES * 2 * MAX_KB * BLK_BYTES;      // Z_STAGES x [hi kb blocks | lo kb blocks]
  float* sCz = reinterpret_cast<float*>(sZ + Z_STAGES * 2 * MAX_KB * BLK_BYTES);     // mpad floats
  float* sXn = sCz + cz_bytes(p.m) / 4;                      // XN_RING x TM
  Barriers* bars = reinterpret_cast<Barriers*>(sXn + XN_RING * TM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = p.nchunks;

  if (tid == 0) {
    for (int s = 0; s < Z_STAGES; ++s) { mbar_init(&bars->z_full[s], 1); mbar_init(&bars->z_empty[s], 1); }
    for (int s = 0; s < X_STAGES; ++s) { mbar_init(&bars->x_full[s], CONV_WARPS); mbar_init(&bars->x_empty[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], EPI_WARPS); }
    mbar_fence_init();
  }
  for (int i = tid; i < mpad(p.m); i += THREADS) sCz[i] = reinterpret_cast<const float*>(p.pack)[i];
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer: inducing chunks, re-streamed (L2-resident) for every row tile =====
    if (lane == 0) {
      const uint8_t* chunks = p.pack + cz_bytes(p.m);
      const uint32_t cbytes = (uint32_t)chunk_bytes(kb);
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % Z_STAGES;
          mbar_wait(&bars->z_empty[s], ((it / Z_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->z_full[s], cbytes);
          uint8_t* dst = sZ + (size_t)s * 2 * MAX_KB * BLK_BYTES;
          const uint8_t* src = chunks + (size_t)c * cbytes;
          for (uint32_t o = 0; o < cbytes; o += BLK_BYTES) bulk_g2s(dst + o, src + o, BLK_BYTES, &bars->z_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(TM, TN);
      uint32_t it = 0, tl = 0;
      for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++tl) {
        const int xs = tl % X_STAGES;
        const uint32_t xa = smem_u32(sX + (size_t)xs * 2 * MAX_KB * BLK_BYTES);
        mbar_wait(&bars->x_full[xs], (tl / X_STAGES) & 1);
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int zs = it % Z_STAGES, as = it % ACC_STAGES;
          mbar_wait(&bars->z_full[zs], (it / Z_STAGES) & 1);
          mbar_wait(&bars->acc_empty[as], ((it / ACC_STAGES) & 1) ^ 1);
          tc_fence_after();
          const uint32_t za = smem_u32(sZ + (size_t)zs * 2 * MAX_KB * BLK_BYTES);
          const uint32_t dcol = tmem_base + (uint32_t)(as * TN);
          uint32_t acc = 0;
          // corrections first (lo x hi, hi x lo), the leading product last
#pragma unroll 1
          for (int combo = 0; combo < 3; ++combo) {
            const uint32_t aoff = (combo == 0) ? (uint32_t)kb * BLK_BYTES : 0u;      // X lo | hi | hi
            const uint32_t boff = (combo == 1) ? (uint32_t)kb * BLK_BYTES : 0u;      // Z hi | lo | hi
            for (int ks = 0; ks < p.ksteps; ++ks) {
              const uint32_t koff = (uint32_t)(ks >> 2) * BLK_BYTES + (uint32_t)(ks & 3) * 32;
              mma_tf32(dcol, smem_desc_sw128(xa + aoff + koff), smem_desc_sw128(za + boff + koff), idesc, acc);
              acc = 1;
            }
          }
          tc_commit(&bars->z_empty[zs]);            // operand buffer free once these MMAs have read it
          tc_commit(&bars->acc_full[as]);           // accumulators complete
        }
        tc_commit(&bars->x_empty[xs]);              // X tile free
      }
    }
    __syncwarp();
  } else if (warp < 2 + CONV_WARPS) {
    // ===== X converters: FP64 rows -> scaled hi / lo TF32 images, row norms =====
    const int cw = warp - 2;
    const int d = p.d;
    const int q0 = 2 * lane;                                     // this lane's feature pair
    const bool have = q0 < d;
    const double il0 = have ? 1.0 / p.ell[q0] : 0.0;
    const double il1 = (q0 + 1 < d) ? 1.0 / p.ell[q0 + 1] : 0.0;
    const bool in_k = q0 < kb * KBLK;
    const uint32_t blk = (uint32_t)(q0 / KBLK) * BLK_BYTES;
    const int kq = q0 % KBLK;
    uint32_t tl = 0;
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++tl) {
      const int xs = tl % X_STAGES;
      uint8_t* sXt = sX + (size_t)xs * 2 * MAX_KB * BLK_BYTES;
      mbar_wait(&bars->x_empty[xs], ((tl / X_STAGES) & 1) ^ 1);
      float* xn = sXn + (tl % XN_RING) * TM;
      if ((p.debug & 2) && tl >= X_STAGES) { __syncwarp(); if (lane == 0) mbar_arrive(&bars->x_full[xs]); continue; }
      const int64_t row0 = t * TM;
      constexpr int RB = 8;                                       // rows in flight per warp
      for (int rb = cw * (TM / CONV_WARPS); rb < (cw + 1) * (TM / CONV_WARPS); rb += RB) {
        double2 v[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int64_t row = row0 + rb + i;
          v[i] = make_double2(0.0, 0.0);
          if (have && row < p.n) v[i] = *reinterpret_cast<const double2*>(p.X + row * p.ldx + q0);
        }
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int r = rb + i;
          const double a = v[i].x * il0, b = v[i].y * il1;           // GPy: X / lengthscale
          double nn = fma(a, a, b * b);
          uint32_t h0, l0, h1, l1;
          split_tf32_bits(a, h0, l0);
          split_tf32_bits(b, h1, l1);
          if (in_k) {
            const uint32_t off = blk + sw128_offset(r, kq);
            *reinterpret_cast<uint2*>(sXt + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(sXt + (size_t)kb * BLK_BYTES + off) = make_uint2(l0, l1);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
          if (lane == 0) xn[r] = (float)(-0.5 * nn * 1.4426950408889634) + p.l2sf2;
        }
      }
      fence_proxy_async();                       // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_full[xs]);
    }
  } else {
    // ===== epilogue: TMEM -> exponent -> exp2 -> FP64 -> HBM =====
    const int ew = warp - 2 - CONV_WARPS;
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int half = ew >> 2;                    // column half of the 128-column stage
    const int row_in_tile = q * 32 + lane;
    const bool vec_ok = (p.ldk % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 31) == 0);
    uint32_t it = 0, tl = 0;
    for (int64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++tl) {
      float cx = 0.f;
      bool have_cx = false;
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int as = it % ACC_STAGES;
        mbar_wait(&bars->acc_full[as], (it / ACC_STAGES) & 1);
        tc_fence_after();
        if (!have_cx) { cx = sXn[(tl % XN_RING) * TM + row_in_tile]; have_cx = true; }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TN + half * 64);
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[as]);         // the stage is free: values are in registers
        const int col0 = c * TN + half * 64;
        if (p.debug & 1) continue;
        // A thread holds 32 consecutive columns of ITS row; stored like that, one warp instruction would
        // touch 32 different 128-byte lines (one 32-byte sector each) and the L1 store path, not HBM,
        // bounds the kernel.  So the four lanes of a quad first exchange 4-column units (a 4 x 4
        // transpose by two shuffle butterflies): afterwards lane j of the quad holds, for each of the
        // quad's four rows, columns [16 h + 4 j, +4) -- and a 32-byte store per lane makes the quad
        // write one full line of one row.
        const int quad_row0 = (int)(t * TM) + q * 32 + (lane & ~3);
        const int jq = lane & 3;
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          uint32_t* v = hb ? v1 : v0;
          float w[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = fmaf(__uint_as_float(v[i]), LOG2E, cx);      // row term first
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // units j = 0..3 of this h: w[16 h + 4 j + e]
#pragma unroll
            for (int bit = 1; bit <= 2; bit <<= 1) {
              const bool up = (lane & bit) != 0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j & bit) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float& lo_u = w[16 * h + 4 * j + e];
                  float& hi_u = w[16 * h + 4 * (j | bit) + e];
                  const float send = up ? lo_u : hi_u;
                  const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
                  if (up) lo_u = recv; else hi_u = recv;
                }
              }
            }
            // now w[16 h + 4 rho + e] = (row quad_row0 + rho, column col0 + 32 hb + 16 h + 4 jq + e)
            const int cc = hb * 32 + h * 16 + jq * 4;
            const float4 z4 = *reinterpret_cast<const float4*>(sCz + col0 + cc);
            const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
            for (int rho = 0; rho < 4; ++rho) {
              const int64_t orow = (int64_t)quad_row0 + rho;
              double o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float ex = w[16 * h + 4 * rho + e] + zz[e];
                o[e] = ex >= p.l2sf2 ? p.sf2 : (double)ex2_approx(ex);          // r^2 clipped at 0 (GPy): exactly sf2
              }
              if (orow < p.n) {
                double* out = p.K + orow * p.ldk + col0 + cc;
                if (vec_ok && col0 + cc + 3 < p.m) {
                  st_v4_f64(out, o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (col0 + cc + e < p.m) out[e] = o[e];
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline size_t smem_bytes(int m) {
  return 1024 /* alignment slack */ + (size_t)2 * MAX_KB * BLK_BYTES * (X_STAGES + Z_STAGES) + cz_bytes(m) +
         (size_t)XN_RING * TM * 4 + sizeof(Barriers) + 16;
}

}  // namespace tf32

// =================================================================================================
// TF32-split posterior-mean gradients from the STORED cross-covariance (units K4 of the path):
//     G_iq = sum_j K_ij c_j (z_jq - x_iq) / l_q^2 = (K B)_i,1+q - (K B)_i,0 x_iq / l_q^2,
//     B_j,0 = c_j,  B_j,1+q = c_j z_jq / l_q^2,   c = alpha * scale
// i.e. ONE tall-skinny contraction K (n x m, FP64 in HBM) times B (m x (d + 1)) on tcgen05, the FP64
// operand converted to hi / lo TF32 on the fly.  Bound by the 8 m bytes per point it reads.
//   warp 0      producer : B k-blocks (32 inducing points: hi | lo images, prepacked) by TMA bulk copy
//   warp 1      MMA      : 12 tcgen05.mma (3 products x 4 k-steps) per k-block into a 128 x NPAD accumulator
//   warps 2-17  convert  : Kfu tile k-block (128 rows x 32 columns FP64, 32-byte loads, three blocks in
//                          flight in registers) -> hi / lo TF32 by a Veltkamp split and integer repacking
//                          (no FP64 <-> FP32 conversion instructions) -> 128-byte-swizzled shared memory
//   warps 18-21 epilogue : TMEM -> FP64, subtract rowsum * x / l^2, store the G row
// =================================================================================================
namespace g32 {
using namespace tf32;

constexpr int STAGES = 4;
constexpr int NPAD_MAX = 80;                         // d <= 64: 1 + 64 columns, rounded up to 16
constexpr int A_BYTES = 2 * BLK_BYTES;               // hi | lo: 32 KB
constexpr int B_BYTES_MAX = 2 * NPAD_MAX * 128;      // hi | lo: 20 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;   // 52 KB (multiple of 1024)
constexpr int CONV_W = 16, EPI_W = 4;
constexpr int PF = 3;                                // k-blocks of Kfu loads in flight per converter thread
constexpr int NTHREADS = 32 * (2 + CONV_W + EPI_W);

struct GParams {
  const double* Kin; int64_t ldk;
  const double* X; int64_t ldx;
  const double* ell;
  const uint8_t* pack;
  double* G; int64_t ldg;
  int64_t n, ntiles;
  int m, d, npad, kblocks;
  double sf2;
};

struct __align__(8) GBars {
  uint64_t full[STAGES], empty[STAGES];
  uint64_t acc_full[2], acc_empty[2];
};

__global__ void pack_grad_tf32_kernel(const double* __restrict__ Z, const double* __restrict__ ell,
                                      const double* __restrict__ coef, double coef_scale, int m, int d, int npad,
                                      uint8_t* __restrict__ pack) {
  const int kblocks = (m + KBLK - 1) / KBLK;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)kblocks * KBLK * npad) return;
  const int j = (int)(idx / npad), nn = (int)(idx % npad);
  double v = 0.0;
  if (j < m) {
    const double c = coef[j] * coef_scale;
    if (nn == 0) v = c;
    else if (nn <= d) { const double l = ell[nn - 1]; v = c * Z[(size_t)j * d + nn - 1] / (l * l); }
  }
  uint32_t hi, lo;
  split_tf32(v, hi, lo);
  uint8_t* blk = pack + (size_t)(j / KBLK) * (2 * npad * 128);
  const uint32_t off = sw128_offset(nn, j % KBLK);
  *reinterpret_cast<uint32_t*>(blk + off) = hi;
  *reinterpret_cast<uint32_t*>(blk + npad * 128 + off) = lo;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) grad_tf32_kernel(const GParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* sIl2 = reinterpret_cast<double*>(smem + (size_t)STAGES * STAGE_BYTES);      // 64 doubles
  GBars* bars = reinterpret_cast<GBars*>(sIl2 + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.kblocks;
  const int64_t my_tiles = blockIdx.x < p.ntiles ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = my_tiles * KB;                       // (tile, k-block) work items of this CTA

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], CONV_W + 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], EPI_W); }
    mbar_fence_init();
  }
  if (tid < 64) { const double l = tid < p.d ? p.ell[tid] : 1.0; sIl2[tid] = tid < p.d ? 1.0 / (l * l) : 0.0; }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bbytes = (uint32_t)(2 * p.npad * 128);

  if (warp == 0) {
    if (lane == 0) {
      for (int64_t it = 0; it < total; ++it) {
        const int s = (int)(it % STAGES), kb = (int)(it % KB);
        mbar_wait(&bars->empty[s], (uint32_t)((it / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[s], bbytes);
        bulk_g2s(smem + (size_t)s * STAGE_BYTES + A_BYTES, p.pack + (size_t)kb * bbytes, bbytes, &bars->full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(TM, p.npad);
      int64_t it = 0;
      for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int slot = (int)(tl & 1);
        mbar_wait(&bars->acc_empty[slot], (uint32_t)((tl >> 1) & 1) ^ 1);
        const uint32_t dcol = tmem_base + (uint32_t)(slot * 128);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          mbar_wait(&bars->full[s], (uint32_t)((it / STAGES) & 1));
          tc_fence_after();
          const uint32_t aa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t ba = aa + A_BYTES;
#pragma unroll 1
          for (int combo = 0; combo < 3; ++combo) {
            const uint32_t aoff = (combo == 0) ? (uint32_t)BLK_BYTES : 0u;            // K lo | hi | hi
            const uint32_t boff = (combo == 1) ? (uint32_t)(p.npad * 128) : 0u;       // B hi | lo | hi
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_tf32(dcol, smem_desc_sw128(aa + aoff + ks * 32), smem_desc_sw128(ba + boff + ks * 32), idesc, acc);
              acc = 1;
            }
          }
          tc_commit(&bars->empty[s]);
        }
        tc_commit(&bars->acc_full[slot]);
      }
    }
    __syncwarp();
  } else if (warp < 2 + CONV_W) {
    // ===== converters: thread = (column group g of 4, rows rbase + 64 i) =====
    // (GPy's dropped pairs -- entries equal to sf2, r = 0 -- are NOT special-cased here: their terms
    //  c_j K (z_j - x_i) cancel between the contraction and the rowsum * x correction to the mode's
    //  rounding level, like every other term)
    const int ct = tid - 64;                          // 0 .. 511
    const int g = ct & 7, rbase = ct >> 3;            // rows rbase + 64 i, i = 0 .. 1
    // fast path: 32-byte aligned rows and m a multiple of 32 -> one unpredicated 32-byte load per item
    const bool fastp = (p.ldk % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Kin) & 31) == 0) && (p.m % KBLK == 0);
    double buf[PF][2][4];
    // (tile, k-block) of the next load, advanced incrementally (no 64-bit divisions in the loop)
    int64_t ld_tile = blockIdx.x;
    int ld_kb = 0;
    const double* rp0 = nullptr;                      // row pointers of the tile being loaded (+ 4 g)
    const double* rp1 = nullptr;
    auto set_tile = [&]() {
      const int64_t r0 = ld_tile * TM + rbase, r1 = r0 + 64;
      rp0 = r0 < p.n ? p.Kin + r0 * p.ldk + 4 * g : nullptr;
      rp1 = r1 < p.n ? p.Kin + r1 * p.ldk + 4 * g : nullptr;
    };
    set_tile();
    auto load1 = [&](const double* rp, double (&dst)[4]) {
      dst[0] = dst[1] = dst[2] = dst[3] = 0.0;
      if (rp != nullptr) {
        const double* src = rp + ld_kb * KBLK;
        if (fastp) {
          asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f64 {%0, %1, %2, %3}, [%4];"
                       : "=d"(dst[0]), "=d"(dst[1]), "=d"(dst[2]), "=d"(dst[3]) : "l"(src));
        } else {
          const int c0 = ld_kb * KBLK + 4 * g;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + e < p.m) dst[e] = __ldg(src + e);
        }
      }
    };
    auto load = [&](double (&dst)[2][4]) {
      load1(rp0, dst[0]);
      load1(rp1, dst[1]);
      if (++ld_kb == KB) { ld_kb = 0; ld_tile += gridDim.x; set_tile(); }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < total) load(buf[u]);
    uint32_t s = 0, ph = 1;                           // stage and the parity to wait for on its empty barrier
    const uint32_t off0 = sw128_offset(rbase, 4 * g), off1 = sw128_offset(rbase + 64, 4 * g);
    for (int64_t it0 = 0; it0 < total; it0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t it = it0 + u;
        if (it < total) {
          mbar_wait(&bars->empty[s], ph);
          uint8_t* sA = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32_bits_nonneg(buf[u][i][e], hi[e], lo[e]);
            const uint32_t off = i ? off1 : off0;
            *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(sA + BLK_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->full[s]);
          if (it + PF < total) load(buf[u]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: a thread per row, 16 features at a time =====
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const bool xvec = (p.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 31) == 0);
    const bool gvec = (p.ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.G) & 31) == 0);
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int slot = (int)(tl & 1);
      const int64_t row = (blockIdx.x + tl * gridDim.x) * TM + row_in_tile;
      const bool rowok = row < p.n;
      const double* xr = p.X + (rowok ? row : 0) * p.ldx;
      double* gr = p.G + (rowok ? row : 0) * p.ldg;
      mbar_wait(&bars->acc_full[slot], (uint32_t)((tl >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128);
      uint32_t rsb;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(rsb) : "r"(taddr) : "memory");
      double rs = 0.0;
#pragma unroll 1
      for (int c16 = 0; c16 < p.d; c16 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + 1 + c16, v);                       // features c16 .. c16 + 15 (column 0 holds the row sum)
        double x[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int c0 = c16 + 4 * q4;
          x[4 * q4] = x[4 * q4 + 1] = x[4 * q4 + 2] = x[4 * q4 + 3] = 0.0;
          if (rowok && c0 < p.d) {
            if (xvec && c0 + 3 < p.d) {
              asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                           : "=d"(x[4 * q4]), "=d"(x[4 * q4 + 1]), "=d"(x[4 * q4 + 2]), "=d"(x[4 * q4 + 3]) : "l"(xr + c0));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c0 + e < p.d) x[4 * q4 + e] = __ldg(xr + c0 + e);
            }
          }
        }
        tmem_ld_wait();
        if (c16 == 0) rs = (double)__uint_as_float(rsb);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int c0 = c16 + 4 * q4;
          double o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = fma(-rs * x[4 * q4 + e], sIl2[(c0 + e) & 63], (double)__uint_as_float(v[4 * q4 + e]));
          if (rowok && c0 < p.d) {
            if (gvec && c0 + 3 < p.d) {
              st_v4_f64(gr + c0, o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c0 + e < p.d) gr[c0 + e] = o[e];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[slot]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

inline size_t smem_bytes() { return 1024 + (size_t)STAGES * STAGE_BYTES + 64 * 8 + sizeof(GBars) + 16; }

}  // namespace g32


size_t pack_tf32_bytes(int m, int d) {
  const int kb = (d + tf32::KBLK - 1) / tf32::KBLK;
  return tf32::cz_bytes(m) + (size_t)(tf32::mpad(m) / tf32::TN) * tf32::chunk_bytes(kb);
}

cudaError_t launch_pack_tf32(const double* Z, const double* ell, int m, int d, void* pack, cudaStream_t st) {
  const int kb = (d + tf32::KBLK - 1) / tf32::KBLK;
  const int warps = tf32::mpad(m);
  tf32::pack_tf32_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(Z, ell, m, d, kb, static_cast<uint8_t*>(pack));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_kuf_tf32(const double* X, int64_t ldx, int64_t n, int d, const double* ell, const void* pack, int m,
                            double sf2, double* K, int64_t ldk, int sms, cudaStream_t st) {
  tf32::Params p{};
  p.X = X; p.ldx = ldx; p.n = n; p.d = d; p.ell = ell; p.pack = static_cast<const uint8_t*>(pack); p.m = m;
  p.nchunks = tf32::mpad(m) / tf32::TN;
  p.kb = (d + tf32::KBLK - 1) / tf32::KBLK;
  p.ksteps = (d + 7) / 8;
  p.l2sf2 = (float)log2(sf2);
  p.sf2 = sf2;
  p.K = K; p.ldk = ldk;
  p.ntiles = (n + tf32::TM - 1) / tf32::TM;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  { static const int dbg = [] { const char* e = getenv("EDRGP_TF32_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }
  const size_t smem = tf32::smem_bytes(m);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(tf32::kuf_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tf32::kuf_tf32_kernel<<<grid, tf32::THREADS, smem, st>>>(p);
  count_launch();
  return cudaGetLastError();
}

// =================================================================================================
// TF32-split weights of the hyper-parameter gradient (optimisation path, unit a7):
//     T = K o (c_ya y alpha^T + K M'),   rowsum(T)        (M' = c_km M, symmetric, m <= 512)
// i.e. the n x m x m contraction U = K M' on tcgen05 -- 2 n m^2 flop, the largest cost of an objective +
// gradient evaluation -- with the FP64 K tile converted on the fly exactly as in grad_tf32_kernel, all
// four 128-column accumulators of a row tile live in the 512 TMEM columns, and an FP64 epilogue that
// re-reads the K tile (L2), forms T and streams it out with the quad-transposed 32-byte accesses of
// kuf_tf32_kernel.  Column sums of T are left to a column-moment pass.
//   warp 0      producer : M' sub-blocks (32 k x 256 columns: hi | lo, 64 KB) by TMA bulk copies, ring of 2
//   warp 1      MMA      : per k-block and 256-column group 12 tcgen05.mma (N = 256: in SS mode every MMA pays
//                          ~170 cycles for fetching its 128 x 8 A slice from shared memory whatever N is --
//                          tools/umma_rate.cu -- so wide MMAs halve that overhead)
//   warps 2-17  convert  : K tile k-blocks -> hi / lo TF32 (two 32 KB stages, three blocks in flight in registers)
//   warps 18-21 epilogue
// =================================================================================================
namespace w32 {
using namespace tf32;

constexpr int SA = 2, SB = 2;
constexpr int A_BYTES = 2 * BLK_BYTES;               // 32 KB
constexpr int NSUB_MAX = 256;                        // output columns per MMA (and per M' sub-block)
constexpr int B_BYTES = 2 * NSUB_MAX * 128;          // 64 KB: 256 columns x 32 k, hi | lo
constexpr int CONV_W = 16, EPI_W = 4, PF = 3;
constexpr int NTHREADS = 32 * (2 + CONV_W + EPI_W);
constexpr int MAX_M = 512;

struct WParams {
  const double* Kin; int64_t ldk;
  const uint8_t* pack;
  const double* y; const double* alpha; double c_ya;
  double* T; int64_t ldt;
  double* rowsum;
  int64_t n, ntiles;
  int m, kblocks, nchunks;     // nchunks: 128-column accumulator chunks (epilogue granularity)
  int nsub, nhalves;           // MMA width (256, or 128 when m <= 128) and the number of such column groups
};

struct __align__(8) WBars {
  uint64_t a_full[SA], a_empty[SA];
  uint64_t b_full[SB], b_empty[SB];
  uint64_t acc_full[4], acc_empty[4];
};

// pack[(kb * nhalves + h)] = hi | lo images of M'[nsub h + r][32 kb + kk] (rows = output columns, K-major)
__global__ void pack_weights_tf32_kernel(const double* __restrict__ M, int64_t ldm, double scale, int m, int nsub,
                                         int nhalves, uint8_t* __restrict__ pack) {
  const int kblocks = (m + KBLK - 1) / KBLK;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)kblocks * nhalves * nsub * KBLK;
  if (idx >= total) return;
  const int kk = (int)(idx % KBLK);
  const int r = (int)((idx / KBLK) % nsub);
  const int h = (int)((idx / ((int64_t)KBLK * nsub)) % nhalves);
  const int kb = (int)(idx / ((int64_t)KBLK * nsub * nhalves));
  const int j = h * nsub + r, k = kb * KBLK + kk;
  double v = 0.0;
  if (j < m && k < m) v = scale * M[(int64_t)j * ldm + k];
  uint32_t hi, lo;
  split_tf32(v, hi, lo);
  uint8_t* blk = pack + (size_t)(kb * nhalves + h) * (2 * nsub * 128);
  const uint32_t off = sw128_offset(r, kk);
  *reinterpret_cast<uint32_t*>(blk + off) = hi;
  *reinterpret_cast<uint32_t*>(blk + nsub * 128 + off) = lo;
}

__global__ void __launch_bounds__(NTHREADS, 1) weights_tf32_kernel(const WParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                           // SA x 32 KB
  uint8_t* sB = smem + (size_t)SA * A_BYTES;                    // SB x 32 KB
  double* sAlpha = reinterpret_cast<double*>(sB + (size_t)SB * B_BYTES);      // MAX_M doubles
  WBars* bars = reinterpret_cast<WBars*>(sAlpha + MAX_M);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.kblocks, NC = p.nchunks, NH = p.nhalves;
  const uint32_t bbytes = (uint32_t)(2 * p.nsub * 128);
  const int64_t my_tiles = blockIdx.x < p.ntiles ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = my_tiles * KB;

  if (tid == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&bars->a_full[s], CONV_W); mbar_init(&bars->a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], EPI_W); }
    mbar_fence_init();
  }
  for (int i = tid; i < MAX_M; i += NTHREADS) sAlpha[i] = (i < p.m && p.alpha != nullptr) ? p.alpha[i] : 0.0;
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 1;
      for (int64_t it = 0; it < total; ++it) {
        const int kb = (int)(it % KB);
        for (int h = 0; h < NH; ++h) {
          mbar_wait(&bars->b_empty[s], ph);
          mbar_arrive_expect_tx(&bars->b_full[s], bbytes);
          const uint8_t* src = p.pack + (size_t)(kb * NH + h) * bbytes;
          for (uint32_t o = 0; o < bbytes; o += 16384)
            bulk_g2s(sB + (size_t)s * B_BYTES + o, src + o, 16384, &bars->b_full[s]);
          if (++s == SB) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(TM, p.nsub);
      const int cph = p.nsub / TN;                       // accumulator chunks per MMA column group
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int64_t tl = 0; tl < my_tiles; ++tl) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&bars->a_full[sa], pa);
          const uint32_t aa = smem_u32(sA + (size_t)sa * A_BYTES);
          for (int h = 0; h < NH; ++h) {
            mbar_wait(&bars->b_full[sb], pb);
            if (kb == 0)
              for (int c = h * cph; c < (h + 1) * cph && c < NC; ++c) mbar_wait(&bars->acc_empty[c], (uint32_t)(tl & 1) ^ 1);
            tc_fence_after();
            const uint32_t ba = smem_u32(sB + (size_t)sb * B_BYTES);
            const uint32_t dcol = tmem_base + (uint32_t)(h * p.nsub);
            uint32_t acc = kb > 0 ? 1u : 0u;
#pragma unroll 1
            for (int combo = 0; combo < 3; ++combo) {
              const uint32_t aoff = (combo == 0) ? (uint32_t)BLK_BYTES : 0u;          // K lo | hi | hi
              const uint32_t boff = (combo == 1) ? (uint32_t)(p.nsub * 128) : 0u;     // M' hi | lo | hi
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                mma_tf32(dcol, smem_desc_sw128(aa + aoff + ks * 32), smem_desc_sw128(ba + boff + ks * 32), idesc, acc);
                acc = 1;
              }
            }
            tc_commit(&bars->b_empty[sb]);
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          tc_commit(&bars->a_empty[sa]);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        for (int c = 0; c < NC; ++c) tc_commit(&bars->acc_full[c]);
      }
    }
    __syncwarp();
  } else if (warp < 2 + CONV_W) {
    const int ct = tid - 64;                          // 0 .. 511
    const int g = ct & 7, rbase = ct >> 3;
    const bool fastp = (p.ldk % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Kin) & 31) == 0) && (p.m % KBLK == 0);
    double buf[PF][2][4];
    int64_t ld_tile = blockIdx.x;
    int ld_kb = 0;
    const double* rp0 = nullptr;
    const double* rp1 = nullptr;
    auto set_tile = [&]() {
      const int64_t r0 = ld_tile * TM + rbase, r1 = r0 + 64;
      rp0 = r0 < p.n ? p.Kin + r0 * p.ldk + 4 * g : nullptr;
      rp1 = r1 < p.n ? p.Kin + r1 * p.ldk + 4 * g : nullptr;
    };
    set_tile();
    auto load1 = [&](const double* rp, double (&dst)[4]) {
      dst[0] = dst[1] = dst[2] = dst[3] = 0.0;
      if (rp != nullptr) {
        const double* src = rp + ld_kb * KBLK;
        if (fastp) {
          asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f64 {%0, %1, %2, %3}, [%4];"
                       : "=d"(dst[0]), "=d"(dst[1]), "=d"(dst[2]), "=d"(dst[3]) : "l"(src));
        } else {
          const int c0 = ld_kb * KBLK + 4 * g;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + e < p.m) dst[e] = __ldg(src + e);
        }
      }
    };
    auto load = [&](double (&dst)[2][4]) {
      load1(rp0, dst[0]);
      load1(rp1, dst[1]);
      if (++ld_kb == KB) { ld_kb = 0; ld_tile += gridDim.x; set_tile(); }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < total) load(buf[u]);
    uint32_t s = 0, ph = 1;
    const uint32_t off0 = sw128_offset(rbase, 4 * g), off1 = sw128_offset(rbase + 64, 4 * g);
    for (int64_t it0 = 0; it0 < total; it0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t it = it0 + u;
        if (it < total) {
          mbar_wait(&bars->a_empty[s], ph);
          uint8_t* dstA = sA + (size_t)s * A_BYTES;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32_bits_nonneg(buf[u][i][e], hi[e], lo[e]);
            const uint32_t off = i ? off1 : off0;
            *reinterpret_cast<uint4*>(dstA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(dstA + BLK_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->a_full[s]);
          if (it + PF < total) load(buf[u]);
          if (++s == SA) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: S = c_ya y alpha^T + U -> the T buffer (quad-transposed 32-byte stores); the product with
    // K and the row sums follow in mul_rowsum_kernel =====
    const int q = warp & 3;
    const int jq = lane & 3;
    const bool vec_ok = (p.ldt % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.T) & 31) == 0);
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int64_t quad_row0 = (blockIdx.x + tl * gridDim.x) * TM + q * 32 + (lane & ~3);
      double yv[4];
#pragma unroll
      for (int rho = 0; rho < 4; ++rho)
        yv[rho] = (p.y != nullptr && quad_row0 + rho < p.n) ? p.c_ya * __ldg(p.y + quad_row0 + rho) : 0.0;
      for (int c = 0; c < NC; ++c) {
        mbar_wait(&bars->acc_full[c], (uint32_t)(tl & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * TN);
#pragma unroll 1
        for (int hb = 0; hb < 4; ++hb) {
          uint32_t v[32];
          tmem_ld32(taddr + hb * 32, v);
          tmem_ld_wait();
          if (hb == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[c]);
          }
          float w[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = __uint_as_float(v[i]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int col0 = c * TN + hb * 32 + h * 16 + jq * 4;
#pragma unroll
            for (int bit = 1; bit <= 2; bit <<= 1) {
              const bool up = (lane & bit) != 0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j & bit) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float& lo_u = w[16 * h + 4 * j + e];
                  float& hi_u = w[16 * h + 4 * (j | bit) + e];
                  const float send = up ? lo_u : hi_u;
                  const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
                  if (up) lo_u = recv; else hi_u = recv;
                }
              }
            }
            // w[16 h + 4 rho + e] = U(row quad_row0 + rho, column col0 + e)
            double al[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) al[e] = sAlpha[(col0 + e) & (MAX_M - 1)];
#pragma unroll
            for (int rho = 0; rho < 4; ++rho) {
              const int64_t row = quad_row0 + rho;
              double o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                o[e] = fma(yv[rho], al[e], (double)w[16 * h + 4 * rho + e]);
              }
              if (row < p.n && p.T != nullptr) {
                double* out = p.T + row * p.ldt + col0;
                if (vec_ok && col0 + 3 < p.m) {
                  st_v4_f64(out, o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (col0 + e < p.m) out[e] = o[e];
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// T <- K o T (T holds S on entry), rowsum_i = sum_j T_ij: one streaming pass, a warp per row.  Kept out of the
// GEMM kernel's epilogue on purpose: there the K entries had to be re-read ~100 us after the converters'
// pass, came from HBM again, and with few loads in flight per epilogue thread that latency -- while all
// 512 TMEM columns blocked the next tile's MMAs -- was three quarters of the kernel's time.
__global__ void __launch_bounds__(256) mul_rowsum_kernel(const double* __restrict__ K, int64_t ldk, double* __restrict__ T,
                                                         int64_t ldt, int64_t n, int m, double* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const double* kr = K + row * ldk;
    double* tr = T + row * ldt;
    double acc = 0.0;
    for (int j = 2 * lane; j < m; j += 64) {
      if (j + 1 < m) {
        const double2 kv = *reinterpret_cast<const double2*>(kr + j);
        double2 tv = *reinterpret_cast<const double2*>(tr + j);
        tv.x *= kv.x; tv.y *= kv.y;
        acc += tv.x + tv.y;
        *reinterpret_cast<double2*>(tr + j) = tv;
      } else {
        const double t = tr[j] * kr[j];
        acc += t;
        tr[j] = t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && rowsum != nullptr) rowsum[row] = acc;
  }
}

inline size_t smem_bytes() {
  return 1024 + (size_t)SA * A_BYTES + (size_t)SB * B_BYTES + MAX_M * 8 + sizeof(WBars) + 16;
}

}  // namespace w32

static int weights_nsub(int m) { return m > tf32::TN ? w32::NSUB_MAX : tf32::TN; }

size_t pack_weights_tf32_bytes(int m) {
  const int kb = (m + tf32::KBLK - 1) / tf32::KBLK, nsub = weights_nsub(m), nh = (m + nsub - 1) / nsub;
  return (size_t)kb * nh * 2 * nsub * 128;
}

cudaError_t launch_pack_weights_tf32(const double* M, int64_t ldm, double scale, int m, void* pack, cudaStream_t st) {
  const int kb = (m + tf32::KBLK - 1) / tf32::KBLK, nsub = weights_nsub(m), nh = (m + nsub - 1) / nsub;
  const int64_t total = (int64_t)kb * nh * nsub * tf32::KBLK;
  w32::pack_weights_tf32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(M, ldm, scale, m, nsub, nh,
                                                                                static_cast<uint8_t*>(pack));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_weights_tf32(const double* K, int64_t n, int m, int64_t ldk, const void* pack, const double* y,
                                const double* alpha, double c_ya, double* T, int64_t ldt, double* rowsum, int sms,
                                cudaStream_t st) {
  w32::WParams p{};
  p.Kin = K; p.ldk = ldk; p.pack = static_cast<const uint8_t*>(pack); p.y = y; p.alpha = alpha; p.c_ya = c_ya;
  p.T = T; p.ldt = ldt; p.rowsum = rowsum; p.n = n; p.ntiles = (n + tf32::TM - 1) / tf32::TM; p.m = m;
  p.kblocks = (m + tf32::KBLK - 1) / tf32::KBLK; p.nchunks = (m + tf32::TN - 1) / tf32::TN;
  p.nsub = weights_nsub(m); p.nhalves = (m + p.nsub - 1) / p.nsub;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  const size_t smem = w32::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(w32::weights_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  w32::weights_tf32_kernel<<<grid, w32::NTHREADS, smem, st>>>(p);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  w32::mul_rowsum_kernel<<<sms * 8, 256, 0, st>>>(K, ldk, T, ldt, n, m, rowsum);
  count_launch();
  return cudaGetLastError();
}

size_t pack_grad_tf32_bytes(int m, int d) {
  const int npad = (d + 1 + 15) / 16 * 16;
  return (size_t)((m + tf32::KBLK - 1) / tf32::KBLK) * 2 * npad * 128;
}

cudaError_t launch_pack_grad_tf32(const double* Z, const double* ell, const double* coef, double coef_scale, int m,
                                  int d, void* pack, cudaStream_t st) {
  const int npad = (d + 1 + 15) / 16 * 16;
  const int64_t total = (int64_t)((m + tf32::KBLK - 1) / tf32::KBLK) * tf32::KBLK * npad;
  g32::pack_grad_tf32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, ell, coef, coef_scale, m, d, npad,
                                                                             static_cast<uint8_t*>(pack));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_grad_tf32(const double* X, int64_t ldx, int64_t n, int d, const double* Kin, int64_t ldk, double sf2,
                             const double* ell, const void* pack, int m, double* G, int64_t ldg, int sms,
                             cudaStream_t st) {
  g32::GParams p{};
  p.Kin = Kin; p.ldk = ldk; p.X = X; p.ldx = ldx; p.ell = ell; p.pack = static_cast<const uint8_t*>(pack);
  p.G = G; p.ldg = ldg; p.n = n; p.ntiles = (n + tf32::TM - 1) / tf32::TM; p.m = m; p.d = d;
  p.npad = (d + 1 + 15) / 16 * 16; p.kblocks = (m + tf32::KBLK - 1) / tf32::KBLK; p.sf2 = sf2;
  const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
  const size_t smem = g32::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(g32::grad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  g32::grad_tf32_kernel<<<grid, g32::NTHREADS, smem, st>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
