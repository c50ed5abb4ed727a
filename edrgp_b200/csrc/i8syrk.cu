// Exact-product INT8 symmetric reduction P = Kfu^T Kfu on the tcgen05 tensor cores (sm_100a), the optional
// `stats='int8x6'` route of unit K2 (edrgp_inducing_stats_i8; GPy: tdot(psi1) in VarDTC.inference,
// edrgp/gp_model/base.py:69).
//
// The FP64 tensor pipe runs this reduction at 0.91 of its 37 TFLOP/s; the INT8 tensor cores of a B200 offer 4.5 POP/s.
// Kfu entries lie in [0, sf2], so with t = K / (4 sf2) + 1 in [1, 1.25] the 52 mantissa bits of t are the fixed-point
// fraction x of K / (4 sf2).  x is ROUNDED to 48 bits and written in balanced radix 256: six SIGNED digits d_0 .. d_5 in
// [-128, 127] with x = sum_s d_s 2^-8(s+1) (d_0 <= 65).  Products of digits are exact in the 32-bit integer
// accumulators of tensor memory, digit pairs (a, b) with the same a + b = g share an accumulator, pairs with
// a + b >= 6 are dropped, and
//     P = 16 sf2^2 sum_g 2^-8(g+2) A_g,       A_g = sum_{a+b=g} D_a^T D_b,   g = 0 .. 5  (21 pairs).
// Rounding and balanced digits make both error terms -- the 2^-49 rounding of x and the dropped pairs (weight 2^-64,
// zero-mean products) -- ZERO-MEAN, so over n rows they grow like sqrt(n) and P comes out at FP64 rounding level
// (1e-15 relative; measured in tests/test_i8_stats_gpu.py).  The first version of this route truncated x to unsigned
// bytes: a one-sided error of 2^-48 per entry that adds up linearly (P at 6e-14), enough to lose positive
// definiteness of Kuu + beta P for optimised low-noise models.
//
//   slice_u8_kernel   Kfu (n x m FP64, row-major) -> slice planes in the layout the MMA reads: per 128-row k-block,
//                     per 128-column block and per slice one 16 KB block of 128 "rows" (inducing columns) x 128 bytes
//                     (data rows = the contraction index), K-major, 128-byte swizzled -- so the reduction kernel moves
//                     operands with plain bulk copies.  The same pass accumulates Kfu^T y and y^T y (per-CTA partials).
//   syrk_i8_kernel    one CTA per (output tile 128 x 128, row split), two passes over its rows (weight groups 0-2, then
//                     3-5: six groups of 128 tensor-memory columns do not fit at once): warp 0 streams slices (the B
//                     tile's per k-block, double buffered; the A tile's one at a time through a ring), warp 1 issues
//                     tcgen05.mma.kind::i8 (M = N = 128, K = 32: 84 MMAs per k-block), warps 2-9 drain the
//                     accumulators every 16 384 rows (128 x 128 x 6 pairs x 16 384 rows < 2^31) into FP64 registers.
//   i8_reduce_kernel  sums the row splits in fixed order, scales, writes P (both triangles), b and y^T y.
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace edrgp {
namespace i8 {

constexpr int S = 6;                       // slices (48 bits)
constexpr int GMAX = 5;                    // slice pairs (a, b) with a + b <= GMAX are formed: 21 pairs in 6 weight groups
constexpr int TM = 128, TN = 128;          // output tile: A columns x B columns
constexpr int KBLK = 128;                  // data rows per k-block = bytes per swizzled operand row
constexpr int ABLK = 128 * 128;            // one slice of a 128-column block for one k-block: 16 KB
constexpr int BBLK = ABLK;
constexpr int A_STAGES = 2;                // 2 x 16 KB + 2 x 6 x 16 KB of B + barriers = 225 KB of the 227 KB a CTA may have
constexpr int B_STAGES = 2;
constexpr int BSL = GMAX + 1;              // slices of the B tile resident per k-block (second pass: all six)
constexpr int DRAIN_KB = 128;              // k-blocks between drains of the int32 accumulators: 6 pairs x 128^2 x 16 384 rows < 2^31
constexpr int EPI_W = 8;                   // two warps per tensor-memory lane quarter: 64 output columns per thread
constexpr int NTHREADS = 32 * (2 + EPI_W);
// Two passes over the CTA's rows, because six weight groups of 128 columns do not fit the 512 tensor-memory columns:
//   pass 0  groups 0, 1, 2  (pairs a + b <= 2: 6 pairs, slices 0 .. 2 of both tiles)
//   pass 1  groups 3, 4, 5  (pairs 3 <= a + b <= 5: 15 pairs, all six slices)
// (Group g weighs ~2^-8g relative to P: stopping at g = 4 -- 15 pairs, measured 2.15 ms per 524 288-row block instead
// of this version's time -- leaves P at 1e-11 and alpha at 7e-9, an order worse than the FP64 reduction's own
// rounding; with g = 5 the result is FP64-equivalent: profiles/r02_int8_slice_syrk_study.txt.)
// An MMA costs ~100-170 cycles for fetching its 128 x 32-byte slice of A from shared memory whatever N is (SS mode;
// tools/umma_rate.cu), so wide tiles (N = 128) with few resident groups beat N = 64 with all groups resident (the first
// version of this kernel: 2.49 ms per 524 288-row block against the FP64 reduction's 3.91 ms incl. everything).
__host__ __device__ constexpr int pass_lo(int pass) { return pass == 0 ? 0 : 3; }
__host__ __device__ constexpr int pass_hi(int pass) { return pass == 0 ? 2 : GMAX; }
constexpr int SLICE_THREADS = 256;

__host__ __device__ inline int mpad(int m) { return (m + 127) / 128 * 128; }

struct __align__(8) Bars {
  uint64_t a_full[A_STAGES], a_empty[A_STAGES];
  uint64_t b_full[B_STAGES], b_empty[B_STAGES];
  uint64_t acc_full, acc_empty;
};

struct Params {
  const uint8_t* planes;
  int njb;                                 // 128-column blocks
  int64_t nkb;                             // k-blocks in all
  int ntiles, nsplit;
  int64_t kb_per_split;
  const int* tiles;                        // (ta, tb) per tile
  double* part;                            // [split][tile][TM * TN]
  int accumulate;                          // add to the partials of earlier row blocks instead of overwriting them
};

// ---- tcgen05 wrappers (the TF32 kernels' idioms, tf32.cu) -----------------------------------------------------------
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
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, dense (mma_sm100_desc: c_format 2 at bit 4,
// a_format at bit 7, b_format at bit 10 -- 0 = unsigned, 1 = signed --, N >> 3 at bit 17, M >> 4 at bit 24)
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

// byte offset of (operand row r, contraction byte k) inside a 128-byte-swizzled K-major block
__host__ __device__ inline uint32_t sw128_byte(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15));
}

// ---------------------------------------------------------------------------------------------------------------
// Slicing pass.  A CTA walks k-blocks (128 data rows); thread = (inducing column j of the current 128-column block,
// half h of the rows); per group of 32 rows it forms one 32-byte sector of every digit plane's swizzled row j.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SLICE_THREADS) slice_u8_kernel(const double* __restrict__ K, int64_t n, int m, int64_t ldk,
                                                                  const double* __restrict__ y, double inv4sf2,
                                                                  uint8_t* __restrict__ planes, int njb, int64_t nkb,
                                                                  double* __restrict__ bpart, int accumulate) {
  const int tid = threadIdx.x, jl = tid & 127, h = tid >> 7;
  const int mp = njb * 128;
  __shared__ double ys[KBLK];
  // per-thread accumulators in shared memory (a thread owns its (half, column) slots: no races, no barriers):
  // acc[h][0][j] = sum K_ij y_i, acc[h][1][j] = sum K_ij^2 -- the diagonal of P, which is taken from this FP64 sum:
  // the dropped digit pair (3, 3) is a sum of squares there, the one place where a dropped term is not zero-mean
  extern __shared__ double acc_sm[];
  double* bsum = acc_sm + (size_t)h * 2 * mp;
  double* dsum = bsum + mp;
  for (int i = tid; i < 4 * mp; i += SLICE_THREADS) acc_sm[i] = 0.0;
  double yy = 0.0;
  for (int64_t kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
    const int64_t row0 = kb * KBLK;
    __syncthreads();
    if (tid < KBLK) {
      const int64_t r = row0 + tid;
      const double v = (y != nullptr && r < n) ? y[r] : 0.0;
      ys[tid] = v;
      yy = fma(v, v, yy);
    }
    __syncthreads();
#pragma unroll 1
    for (int jb = 0; jb < njb; ++jb) {
      const int j = jb * 128 + jl;
      const bool jok = j < m;
      uint8_t* blk = planes + (((size_t)kb * njb + jb) * S) * ABLK;
      double bj = 0.0, dj = 0.0;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {                       // 32-row pieces of this thread's half: chunks 2 cp, 2 cp + 1
        const int cp = 2 * h + c;
        // both 16-byte chunks of a 32-byte sector are formed by one thread and stored back to back: 16-byte stores
        // issued 16 loads apart left half-written sectors for L2 to evict (measured: 1.6 x the slice bytes written
        // and 1.1 GB of fill reads per 524 288-row block)
        uint32_t w[S][8];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int i = 0; i < 8; ++i) w[s][i] = 0u;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int il = cp * 32 + e;
          const int64_t r = row0 + il;
          double v = 0.0;
          if (jok && r < n) v = __ldg(K + r * ldk + j);
          bj = fma(v, ys[il], bj);
          dj = fma(v, v, dj);
          const double t = fma(v, inv4sf2, 1.0);
          // 52-bit fraction -> rounded to 48 bits -> balanced digits: adding 0x80 to the five low bytes lets the
          // carries run, after which byte ^ 0x80 (as int8) is the digit; the top byte takes the last carry as it is
          const unsigned long long F = ((unsigned long long)((uint32_t)__double2hiint(t) & 0xFFFFFu) << 32) |
                                       (uint32_t)__double2loint(t);
          // (round half to EVEN: half-up is biased by 2^-53 per entry, which adds up linearly over the rows -- measured
          // 4e-15 on P at 524 288 rows against 1e-16-level noise otherwise)
          const unsigned long long R = (((F + 7ull + ((F >> 4) & 1ull)) >> 4) + 0x0000008080808080ull) ^ 0x0000008080808080ull;
          const uint32_t rl = (uint32_t)R, rh = (uint32_t)(R >> 32);
          const uint32_t q0 = (rh >> 8) & 0xFFu, q1 = rh & 0xFFu, q2 = rl >> 24;
          const uint32_t q3 = (rl >> 16) & 0xFFu, q4 = (rl >> 8) & 0xFFu, q5 = rl & 0xFFu;
          const int wi = e >> 2, sh = (e & 3) * 8;
          w[0][wi] |= q0 << sh; w[1][wi] |= q1 << sh; w[2][wi] |= q2 << sh;
          w[3][wi] |= q3 << sh; w[4][wi] |= q4 << sh; w[5][wi] |= q5 << sh;
        }
        const uint32_t off0 = sw128_byte(jl, cp * 32), off1 = sw128_byte(jl, cp * 32 + 16);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          *reinterpret_cast<uint4*>(blk + (size_t)s * ABLK + off0) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
          *reinterpret_cast<uint4*>(blk + (size_t)s * ABLK + off1) = make_uint4(w[s][4], w[s][5], w[s][6], w[s][7]);
        }
      }
      bsum[j] += bj;
      dsum[j] += dj;
    }
  }
  // per-CTA partials: bpart[cta] = [Kfu^T y (mp) | diag(Kfu^T Kfu) (mp) | y^T y] over this CTA's rows
  double* out = bpart + (size_t)blockIdx.x * (2 * mp + 1);
  __syncthreads();
  for (int i = tid; i < 2 * mp; i += SLICE_THREADS)
    out[i] = acc_sm[i] + acc_sm[2 * mp + i] + (accumulate ? out[i] : 0.0);
  __syncthreads();
  ys[tid & 127] = 0.0;
  __syncthreads();
  if (tid < KBLK) ys[tid] = yy;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < KBLK; ++i) s += ys[i];
    out[2 * mp] = s + (accumulate ? out[2 * mp] : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The reduction proper
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t SMEM_A = (size_t)A_STAGES * ABLK;
constexpr size_t SMEM_B = (size_t)B_STAGES * BSL * BBLK;
inline size_t smem_bytes() { return 1024 + SMEM_A + SMEM_B + sizeof(Bars) + 16; }

__global__ void __launch_bounds__(NTHREADS, 1) syrk_i8_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SMEM_A;
  Bars* bars = reinterpret_cast<Bars*>(smem + SMEM_A + SMEM_B);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % p.ntiles, split = blockIdx.x / p.ntiles;
  const int ta = p.tiles[2 * tile], tb = p.tiles[2 * tile + 1];
  const int64_t kb0 = (int64_t)split * p.kb_per_split;
  int64_t kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.nkb) kb1 = p.nkb;
  const int64_t nk = kb1 > kb0 ? kb1 - kb0 : 0;
  const int64_t ndrain = (nk + DRAIN_KB - 1) / DRAIN_KB;

  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&bars->a_full[s], 1); mbar_init(&bars->a_empty[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1); }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, EPI_W);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      int64_t ia = 0, ib = 0;
      for (int pass = 0; pass < 2; ++pass) {
        const int nsl = pass_hi(pass) + 1;                 // slices 0 .. nsl - 1 of both tiles take part in this pass
        for (int64_t k = 0; k < nk; ++k, ++ib) {
          const int64_t kb = kb0 + k;
          const int bs = (int)(ib % B_STAGES);
          mbar_wait(&bars->b_empty[bs], (uint32_t)((ib / B_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars->b_full[bs], (uint32_t)(nsl * BBLK));
          const uint8_t* bsrc = p.planes + (((size_t)kb * p.njb + tb) * S) * ABLK;
          for (int sl = 0; sl < nsl; ++sl)
            bulk_g2s(sB + ((size_t)bs * BSL + sl) * BBLK, bsrc + (size_t)sl * ABLK, (uint32_t)BBLK, &bars->b_full[bs]);
          const uint8_t* asrc = p.planes + (((size_t)kb * p.njb + ta) * S) * ABLK;
          for (int a = 0; a < nsl; ++a, ++ia) {
            const int as = (int)(ia % A_STAGES);
            mbar_wait(&bars->a_empty[as], (uint32_t)((ia / A_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars->a_full[as], (uint32_t)ABLK);
            bulk_g2s(sA + (size_t)as * ABLK, asrc + (size_t)a * ABLK, (uint32_t)ABLK, &bars->a_full[as]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(TM, TN);
      int64_t ia = 0, ib = 0, idr = 0;                     // ring positions and the number of drains requested so far
      for (int pass = 0; pass < 2; ++pass) {
        const int glo = pass_lo(pass), ghi = pass_hi(pass);
        for (int64_t k = 0; k < nk; ++k, ++ib) {
          const bool first = (k % DRAIN_KB) == 0;
          if (first && idr > 0) {
            // the accumulators of the previous 16 384 rows (or of the previous pass) must have been drained
            mbar_wait(&bars->acc_empty, (uint32_t)((idr - 1) & 1));
            tc_fence_after();
          }
          const int bs = (int)(ib % B_STAGES);
          mbar_wait(&bars->b_full[bs], (uint32_t)((ib / B_STAGES) & 1));
          const uint32_t ba = smem_u32(sB + (size_t)bs * BSL * BBLK);
          for (int a = 0; a <= ghi; ++a, ++ia) {
            const int as = (int)(ia % A_STAGES);
            mbar_wait(&bars->a_full[as], (uint32_t)((ia / A_STAGES) & 1));
            tc_fence_after();
            const uint32_t aa = smem_u32(sA + (size_t)as * ABLK);
            const int blo = glo - a > 0 ? glo - a : 0;
#pragma unroll 1
            for (int b = blo; b <= ghi - a; ++b) {
              const int g = a + b;
              const uint32_t dcol = tmem_base + (uint32_t)((g - glo) * TN);
              // the first pair of a group in a drain interval overwrites: that is the pair with a = 0 (b = g) for the
              // groups g <= ghi of the pass -- a = 0 is issued first and reaches every group of the pass
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_i8(dcol, smem_desc_sw128(aa + ks * 32), smem_desc_sw128(ba + (uint32_t)(b * BBLK) + ks * 32), idesc,
                       (first && a == 0 && ks == 0) ? 0u : 1u);
            }
            tc_commit(&bars->a_empty[as]);
          }
          tc_commit(&bars->b_empty[bs]);
          if ((k % DRAIN_KB) == DRAIN_KB - 1 || k == nk - 1) { tc_commit(&bars->acc_full); ++idr; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: a thread per output row (A column) and column half, 64 FP64 sums in registers =====
    const int q = warp & 3, ch = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    double sum[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) sum[c] = 0.0;
    int64_t idr = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int glo = pass_lo(pass), ghi = pass_hi(pass);
      for (int64_t dr = 0; dr < ndrain; ++dr, ++idr) {
        mbar_wait(&bars->acc_full, (uint32_t)(idr & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 64);
#pragma unroll 1
        for (int g = glo; g <= ghi; ++g) {
          const double w = __hiloint2double((1023 - 8 * (g + 2)) << 20, 0);          // 2^-8(g+2)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)((g - glo) * TN + hh * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[hh * 32 + c] = fma((double)(int)v[c], w, sum[hh * 32 + c]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty);
      }
    }
    double* out = p.part + ((size_t)split * p.ntiles + tile) * (TM * TN) + (size_t)r * TN + ch * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 2) {
      double2 o = make_double2(sum[c], sum[c + 1]);
      if (p.accumulate) { const double2 old = *reinterpret_cast<const double2*>(out + c); o.x += old.x; o.y += old.y; }
      *reinterpret_cast<double2*>(out + c) = o;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// P (+)= scale * sum over splits; b_yy (+)= sum over the slicer's CTAs
__global__ void __launch_bounds__(256) i8_reduce_kernel(const double* __restrict__ part, int nsplit, int ntiles,
                                                        const int* __restrict__ tiles, int m, double scale,
                                                        double* __restrict__ P, int64_t ldp, int accumulate,
                                                        const double* __restrict__ bpart, int nb, int mp,
                                                        double* __restrict__ b_yy) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)ntiles * TM * TN;
  if (idx < total) {
    const int tile = (int)(idx / (TM * TN)), e = (int)(idx % (TM * TN));
    const int ta = tiles[2 * tile], tb = tiles[2 * tile + 1];
    const int j = ta * TM + e / TN, jp = tb * TN + e % TN;
    if (j < m && jp < m) {
      double s = 0.0;
      if (j == jp) {             // the diagonal: the slicer's FP64 sums of squares
        for (int c = 0; c < nb; ++c) s += bpart[(size_t)c * (2 * mp + 1) + mp + j];
      } else {
        for (int sp = 0; sp < nsplit; ++sp) s += part[((size_t)sp * ntiles + tile) * (TM * TN) + e];
        s *= scale;
      }
      P[(int64_t)j * ldp + jp] = accumulate ? P[(int64_t)j * ldp + jp] + s : s;
      // the mirror entry is written here unless the tile is a diagonal one (which computes both triangles itself)
      if (tb > ta) P[(int64_t)jp * ldp + j] = accumulate ? P[(int64_t)jp * ldp + j] + s : s;
    }
  }
  if (b_yy != nullptr && idx <= m) {
    const int col = idx < m ? (int)idx : 2 * mp;
    double s = 0.0;
    for (int c = 0; c < nb; ++c) s += bpart[(size_t)c * (2 * mp + 1) + col];
    b_yy[idx] = accumulate ? b_yy[idx] + s : s;
  }
}

static int tile_list(int m, int* out) {                 // (ta, tb), tb >= ta: the upper triangle of 128 x 128 tiles
  const int na = mpad(m) / TM, nb = na;
  int cnt = 0;
  for (int ta = 0; ta < na; ++ta)
    for (int tb = ta; tb < nb; ++tb) {
      if (out) { out[2 * cnt] = ta; out[2 * cnt + 1] = tb; }
      ++cnt;
    }
  return cnt;
}

}  // namespace i8

// workspace: split partials | slicer partials | tile table | slice planes of one row block (the only part that
// depends on n, so it comes last: the partials of successive row blocks of different sizes share their place)
namespace i8 {
struct Layout {
  int ntiles, nsplit, njb, mp, sgrid;
  size_t part, bpart, tiles, planes, total;
};
static Layout layout(int64_t n, int m, int sms) {
  Layout L;
  L.ntiles = tile_list(m, nullptr);
  L.nsplit = sms / L.ntiles < 1 ? 1 : sms / L.ntiles;
  L.njb = mpad(m) / 128;
  L.mp = L.njb * 128;
  L.sgrid = 2 * sms;
  const int64_t nkb = (n + KBLK - 1) / KBLK;
  size_t o = 0;
  L.part = o;   o += (size_t)L.nsplit * L.ntiles * TM * TN * 8;
  L.bpart = o;  o += (size_t)L.sgrid * (2 * L.mp + 1) * 8;
  L.tiles = o;  o += ((size_t)2 * L.ntiles * 4 + 1023) / 1024 * 1024;
  L.planes = o; o += (size_t)nkb * L.njb * S * ABLK;
  L.total = (o + 1023) / 1024 * 1024;
  return L;
}
}  // namespace i8

size_t inducing_stats_i8_workspace_bytes(int64_t n, int m, int sms) { return i8::layout(n, m, sms).total; }

// One row block: slices, then the reduction into the split partials (first = 0: on top of the earlier blocks').
cudaError_t launch_i8_block(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2, int first,
                            void* workspace, int sms, cudaStream_t st) {
  if (m > 2048) return cudaErrorInvalidValue;
  const i8::Layout L = i8::layout(n, m, sms);
  const int64_t nkb = (n + i8::KBLK - 1) / i8::KBLK;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* tiles = reinterpret_cast<int*>(ws + L.tiles);
  cudaError_t e;
  if (first) {
    int host_tiles[2 * 16 * 17 / 2 * 2];
    i8::tile_list(m, host_tiles);
    if ((e = cudaMemcpyAsync(tiles, host_tiles, (size_t)2 * L.ntiles * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return e;
  }
  const size_t slice_smem = (size_t)4 * L.mp * sizeof(double);
  static bool slice_attr_set = false;
  if (!slice_attr_set && slice_smem > 40 * 1024) {
    if ((e = cudaFuncSetAttribute(i8::slice_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 8)) != cudaSuccess) return e;
    slice_attr_set = true;
  }
  i8::slice_u8_kernel<<<L.sgrid, i8::SLICE_THREADS, slice_smem, st>>>(Kfu, n, m, ldk, y, 0.25 / sf2, ws + L.planes, L.njb, nkb,
                                                             reinterpret_cast<double*>(ws + L.bpart), first ? 0 : 1);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  i8::Params p{};
  p.planes = ws + L.planes; p.njb = L.njb; p.nkb = nkb; p.ntiles = L.ntiles; p.nsplit = L.nsplit;
  p.kb_per_split = (nkb + L.nsplit - 1) / L.nsplit;
  p.tiles = tiles; p.part = reinterpret_cast<double*>(ws + L.part); p.accumulate = first ? 0 : 1;
  const size_t smem = i8::smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(i8::syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    attr_set = true;
  }
  i8::syrk_i8_kernel<<<L.ntiles * L.nsplit, i8::NTHREADS, smem, st>>>(p);
  count_launch();
  return cudaGetLastError();
}

// P (+)= 16 sf2^2 x the sum of the split partials (fixed order), b_yy (+)= the slicer's partials
cudaError_t launch_i8_finish(int m, double sf2, int with_y, double* P, int64_t ldp, double* b_yy, int accumulate,
                             void* workspace, int sms, cudaStream_t st) {
  const i8::Layout L = i8::layout(0, m, sms);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int64_t total = (int64_t)L.ntiles * i8::TM * i8::TN;
  i8::i8_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const double*>(ws + L.part), L.nsplit, L.ntiles, reinterpret_cast<const int*>(ws + L.tiles), m,
      16.0 * sf2 * sf2, P, ldp, accumulate, reinterpret_cast<const double*>(ws + L.bpart), L.sgrid, L.mp,
      with_y ? b_yy : nullptr);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_inducing_stats_i8(const double* Kfu, int64_t n, int m, int64_t ldk, const double* y, double sf2,
                                     double* P, int64_t ldp, double* b_yy, int accumulate, void* workspace, int sms,
                                     cudaStream_t st) {
  cudaError_t e = launch_i8_block(Kfu, n, m, ldk, y, sf2, 1, workspace, sms, st);
  if (e != cudaSuccess) return e;
  return launch_i8_finish(m, sf2, y != nullptr, P, ldp, b_yy, accumulate, workspace, sms, st);
}

}  // namespace edrgp
