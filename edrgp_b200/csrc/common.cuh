// Shared device helpers for the edrgp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edrgp {

constexpr int MT = 32;               // inducing points per streamed tile
constexpr int WARPS = 8;             // compute warps per CTA
constexpr int ROWS_PER_WARP = 16;    // two m8 blocks
constexpr int BM = WARPS * ROWS_PER_WARP;   // 128 data rows per CTA tile

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }
// padded feature count: the k-permutation of the distance GEMM works on groups of 16 features
__host__ __device__ constexpr int padded_dim(int d) { return round_up(d, 16); }
// shared-memory row stride (doubles): == 2 (mod 16) makes both DMMA fragment patterns
// (rows g, cols {0,1,8,9}+2s  and  rows 2t+s, cols g) bank-conflict free for 64-bit loads
__host__ __device__ constexpr int row_stride(int dp) { return dp + 2; }
// doubles per packed inducing tile: MT rows + MT x hz + MT x coef
__host__ __device__ constexpr int pack_tile_doubles(int dp) { return MT * (row_stride(dp) + 2); }

// ---- FP64 tensor core: D(8x8) += A(8x4) * B(4x8); SASS DMMA.8x8x4 ------------------------------
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS UBLKCP) -----------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// order generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- exp(x) for x <= 0 -------------------------------------------------------------------------
// x = n ln2/64 + r, |r| <= ln2/128:  exp(x) = 2^(n >> 6) * T[n & 63] * (1 + r + r^2/2 + ... + r^5/120),
// T[j] = 2^(j/64) from a shared-memory table (exp_table_init).  About 10 FP64 instructions against ~20 for
// libm's exp: the FP64 pipe is shared with the DMMA contraction, so each one counts.  Truncation error
// r^6/720 < 4e-17; results below the normal range (x < -708) flush to zero.
// The table index is data dependent, so a plain 64-entry table (512 B = 4 rows of the 32 banks) made the 16
// lanes of a half-warp -- the unit an 8-byte shared-memory load is served in -- collide on bank pairs at
// random: 56.7 M bank conflicts per 524 288-row launch of kuf_kernel, 2.7 wavefronts per load instead of 1
// (ncu r01).  The table is therefore REPLICATED once per lane of a half-warp: entry j of copy c = lane & 15
// sits at double index 16 j + c, i.e. in bank pair c whatever j is, so a load never conflicts (8 KB).
// `scale` (the kernel variance in kuf_kernel) is folded into the table, so that scale * exp(x) costs no extra
// multiplication and exp(0) * scale is exactly `scale` (T[0] = scale, p = 0).
constexpr int EXP_TABLE_DOUBLES = 64 * 16;
__device__ __forceinline__ void exp_table_init(double* tab, int tid, int nthreads, double scale = 1.0) {
  for (int i = tid; i < EXP_TABLE_DOUBLES; i += nthreads) tab[i] = scale * exp2((double)(i >> 4) * (1.0 / 64.0));
}
// tab_lane = table base + (lane & 15): the caller hoists the lane offset out of its loops
__device__ __forceinline__ double exp_neg(double x, const double* tab_lane) {
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);       // x * 64/ln2 + 1.5 * 2^52
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0x1.62e42fee00000p-7, x);                          // ln2/64, high part (31 bits)
  r = fma(nf, -0x1.a39ef35793c76p-39, r);                               //         low part
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;
  const double tj = tab_lane[(n & 63) << 4];
  const double v = fma(tj, p, tj);
  const int hi = __double2hiint(v) + ((n >> 6) << 20);
  return x < -708.0 ? 0.0 : __hiloint2double(hi, __double2loint(v));
}
// scale * exp(min(x, 0)) with the scale folded into the table; the clip and the underflow test are integer
// compares on the high word (the FP64 pipe, shared with the DMMAs, sees only the 10 arithmetic instructions).
// Underflow is decided on the patched exponent itself (result below the normal range -> 0, as exp_neg does for
// x < -708 at scale 1), so any positive scale below 2^1000 works; x < -1024 is cut off before its n overflows.
__device__ __forceinline__ double exp_clip_scaled(double x, const double* tab_lane) {
  const int xh = __double2hiint(x);
  x = xh < 0 ? x : 0.0;                                                  // min(x, 0): positive sign bit -> 0
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0x1.62e42fee00000p-7, x);
  r = fma(nf, -0x1.a39ef35793c76p-39, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = p * r;
  const double tj = tab_lane[(n & 63) << 4];
  const double v = fma(tj, p, tj);
  const int hi = __double2hiint(v) + ((n >> 6) << 20);
  // (high word of -1024.0 is 0xC0900000; for negative doubles a larger magnitude is a larger word)
  return (hi < 0x00100000 || (unsigned)xh > 0xC0900000u) ? 0.0 : __hiloint2double(hi, __double2loint(v));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace edrgp
