// Small dense FP64 linear algebra on one GPU for the m x m inducing system and the d x d EDR
// matrix: blocked Cholesky, blocked triangular solves, cyclic Jacobi eigensolver.
//
// These stand in for the LAPACK calls GPy makes inside VarDTC.inference (dpotrf via jitchol,
// dtrtrs, dpotri) and for np.linalg.svd in SVDTransformer.fit (edrgp/utils.py:140).  Sizes are
// m <= a few thousand and d <= a few hundred: ~m^3 flops against the ~n m^2 of the statistics
// pass, so the kernels are written for clarity and determinism (32 x 32 blocks, no atomics), not
// for the last TFLOP.  All matrices are row-major.
#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int NBK = 32;

// Cheap FP64 reciprocal / reciprocal square root: FP32 hardware estimate + two Newton steps (relative
// error ~2e-16, i.e. the last bit).  The IEEE divide / sqrt sequences (~40 dependent FP64 instructions
// each) sit on the critical path of the one-warp Cholesky steps and of the Jacobi rotations (where the
// ANGLE only steers convergence -- what must be exact is c^2 + s^2 = 1, which c = rsqrt(1 + t^2),
// s = t c gives).
__device__ __forceinline__ double fast_rsqrt(double x) {      // 1e-30 < x < 1e30
  double y = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double fast_rcp(double x) {        // 1e-30 < |x| < 1e30
  double r = (double)__frcp_rn((float)x);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}


// ---------------------------------------------------------------------------------------------
// generic strided small GEMM:  C[i][j] = beta C[i][j] + alpha sum_k A(i,k) B(k,j)
//   A(i,k) = A[i*sai + k*sak],  B(k,j) = B[k*sbk + j*sbj],  C row-major with ldc.
// lower_only: skip 32x32 output blocks strictly above the diagonal (SYRK-style trailing updates).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_small_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                         int64_t sai, int64_t sak, const double* __restrict__ B,
                                                         int64_t sbk, int64_t sbj, double beta, double* __restrict__ C,
                                                         int64_t ldc, int lower_only) {
  __shared__ double As[NBK][NBK + 1];
  __shared__ double Bs[NBK][NBK + 1];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (lower_only && bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty in 0..7: rows ty, ty+8, ty+16, ty+24
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < K; k0 += NBK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int li = ty + 8 * r;
      // As[li][tx] = A(i0+li, k0+tx);  Bs[li][tx] = B(k0+li, j0+tx)
      const int i = bi * NBK + li, k = k0 + tx;
      As[li][tx] = (i < M && k < K) ? A[i * sai + k * sak] : 0.0;
      const int kk = k0 + li, j = bj * NBK + tx;
      Bs[li][tx] = (kk < K && j < N) ? B[kk * sbk + j * sbj] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < NBK; ++k) {
      const double b = Bs[k][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fma(As[ty + 8 * r][k], b, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = bi * NBK + ty + 8 * r, j = bj * NBK + tx;
    if (i < M && j < N) {
      double* c = C + i * ldc + j;
      *c = (beta == 0.0 ? 0.0 : beta * *c) + alpha * acc[r];
    }
  }
}

static cudaError_t gemm_small(int M, int N, int K, double alpha, const double* A, int64_t sai, int64_t sak,
                              const double* B, int64_t sbk, int64_t sbj, double beta, double* C, int64_t ldc,
                              int lower_only, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + NBK - 1) / NBK, (M + NBK - 1) / NBK);
  gemm_small_kernel<<<grid, 256, 0, st>>>(M, N, K, alpha, A, sai, sak, B, sbk, sbj, beta, C, ldc, lower_only); count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Cholesky, lower, in place (upper triangle left untouched).  info[0] = 1 + column of the first
// non-positive pivot (0 = success), like LAPACK dpotrf.
//
// One blocked step = two launches: potrf_step_kernel factors the 32 x 32 diagonal block in
// registers (lane i holds row i; pivots and column entries travel by warp shuffles) and solves the
// panel below it, every one-warp CTA re-deriving the small factor for itself instead of waiting for
// another launch; gemm_small_kernel then applies the trailing update.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void potf2_regs(double (&a)[NBK], int lane, int nb, int k0, int* info) {
#pragma unroll
  for (int j = 0; j < NBK; ++j) {
    const double djj = __shfl_sync(0xffffffffu, a[j], j);
    if (j < nb && !(djj > 0.0)) {
      if (info != nullptr && lane == 0) atomicCAS(info, 0, k0 + j + 1);
    }
    // pivot through the cheap reciprocal square root when it is in the estimate's range
    const bool fast = djj > 1e-30 && djj < 1e30;
    const double inv = fast ? fast_rsqrt(djj) : 1.0 / sqrt(djj);
    const double ljj = fast ? djj * inv : sqrt(djj);
    const double lij = lane == j ? ljj : a[j] * inv;
    if (lane >= j) a[j] = lij;
#pragma unroll
    for (int c = j + 1; c < NBK; ++c) {
      const double lcj = __shfl_sync(0xffffffffu, a[j], c);
      if (lane >= c) a[c] = fma(-lij, lcj, a[c]);
    }
  }
}

// CTA 0 also writes the factored diagonal block back; CTA b solves rows k0+nb+32b.. of the panel:
// A21 <- A21 L11^-T.
__global__ void __launch_bounds__(32) potrf_step_kernel(double* __restrict__ A, int64_t ld, int k0, int nb, int m,
                                                        int* info) {
  __shared__ double S[NBK][NBK + 1];
  __shared__ double X[NBK][NBK + 1];
  const int lane = threadIdx.x;
  // coalesced load of the diagonal block (lane = column), padded with the identity
  for (int r = 0; r < NBK; ++r) {
    double v = (r == lane) ? 1.0 : 0.0;
    if (r < nb && lane < nb && lane <= r) v = A[(int64_t)(k0 + r) * ld + k0 + lane];
    S[r][lane] = v;
  }
  __syncwarp();
  double a[NBK];
#pragma unroll
  for (int c = 0; c < NBK; ++c) a[c] = S[lane][c];
  potf2_regs(a, lane, nb, k0, blockIdx.x == 0 ? info : nullptr);
#pragma unroll
  for (int c = 0; c < NBK; ++c) S[lane][c] = a[c];          // S = L11 (lower; junk above the diagonal unused)
  __syncwarp();
  if (blockIdx.x == 0) {
    for (int r = 0; r < nb; ++r)
      if (lane <= r && lane < nb) A[(int64_t)(k0 + r) * ld + k0 + lane] = S[r][lane];
  }
  const int row0 = k0 + nb + blockIdx.x * NBK;
  if (row0 >= m) return;
  for (int r = 0; r < NBK; ++r) {
    const int rr = row0 + r;
    X[r][lane] = (rr < m && lane < nb) ? A[(int64_t)rr * ld + k0 + lane] : 0.0;
  }
  __syncwarp();
  {
    // lane = row of the panel: forward substitution against L11^T, entries kept in registers
    double x[NBK];
#pragma unroll
    for (int c = 0; c < NBK; ++c) x[c] = X[lane][c];
    const double dinv = 1.0 / S[lane][lane];          // lane c holds 1 / L11[c][c]
#pragma unroll
    for (int c = 0; c < NBK; ++c) {
      double v = x[c];
#pragma unroll
      for (int q = 0; q < c; ++q) v = fma(-x[q], S[c][q], v);
      x[c] = v * __shfl_sync(0xffffffffu, dinv, c);
    }
#pragma unroll
    for (int c = 0; c < NBK; ++c) X[lane][c] = x[c];
  }
  __syncwarp();
  for (int r = 0; r < NBK; ++r) {
    const int rr = row0 + r;
    if (rr < m && lane < nb) A[(int64_t)rr * ld + k0 + lane] = X[r][lane];
  }
}

cudaError_t launch_potrf(double* A, int m, int64_t ld, int* info, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  for (int k0 = 0; k0 < m; k0 += NBK) {
    const int nb = min(NBK, m - k0);
    const int rest = m - k0 - nb;
    const int ctas = rest > 0 ? (rest + NBK - 1) / NBK : 1;
    potrf_step_kernel<<<ctas, 32, 0, st>>>(A, ld, k0, nb, m, info); count_launch();
    if (rest > 0) {
      // A22 -= L21 L21^T (lower blocks only)
      double* A22 = A + (int64_t)(k0 + nb) * ld + k0 + nb;
      const double* L21 = A + (int64_t)(k0 + nb) * ld + k0;
      e = gemm_small(rest, rest, nb, -1.0, L21, ld, 1, L21, 1, ld, 1.0, A22, ld, 1, st);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Cholesky solve with ONE right-hand side (LAPACK dposv, nrhs = 1), the posterior weights of the
// fixed-hyper-parameter sweep.  What bounds it is the chain of dependent blocked steps, not work, so
// one blocked step is ONE launch: CTA (i, j) of step k re-derives the factor of the 32 x 32 diagonal
// block for itself (a warp; the finished column is broadcast through shared memory), solves the two
// panel tiles it needs (a warp each, a lane per row, right-looking so that the 31 updates behind a
// finished unknown are independent) and applies its own 32 x 32 trailing update.  The factor goes to
// a SEPARATE output matrix -- every CTA of the step reads block column k of A while one of them
// produces L's -- and the right-hand side rides along as one more row of the matrix, which makes the
// forward substitution free: after the last step that row holds L^-1 rhs.  A's lower triangle is
// destroyed; the backward substitution is trsv_inv_kernel (trsv_kernel beyond 704 unknowns).
// ---------------------------------------------------------------------------------------------
// FP64 operations cost ~30-37 cycles each on a dependent chain here (tools/chol_profile.py), so the
// routines below are arranged around the number of DEPENDENT FP64 operations per column.
//
// reciprocal / reciprocal square root of a positive normal double: hardware FP64 seed (~20 bits) + two
// Newton steps, branch free over the whole normal range (the unrolled factorisation below must stay ONE
// basic block for the scheduler to interleave its independent chains)
__device__ __forceinline__ double rcp_seeded(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double rsqrt_seeded(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double hd = 0.5 * d;
  y = y * fma(-hd * y, y, 1.5);
  y = y * fma(-hd * y, y, 1.5);
  return y;
}

// Factor the 32 x 32 block whose row `lane` is in a[] (identity padded).  Column j is broadcast UNSCALED
// through shared memory (col: 64 doubles owned by this warp, alternating halves) as soon as it is final,
// the trailing update uses u_i u_c / d_j with the reciprocal of the pivot, and the next pivot is formed
// on its own lane from that lane's own entry and fetched by one shuffle: the chain from pivot to pivot is
// reciprocal -> fma -> shuffle; the reciprocal square root that scales the column into L runs beside it.
// Returns the reciprocal of this lane's diagonal entry of L; *bad = 1 + index of the first non-positive pivot (0: none).
__device__ __forceinline__ double potf2_bcast(double (&a)[NBK], double* col, int lane, int nb, int k0, int* bad) {
  double dj = __shfl_sync(0xffffffffu, a[0], 0);
  double myinv = 1.0;
  int first_bad = 0;
#pragma unroll
  for (int j = 0; j < NBK; ++j) {
    const double u = a[j];
    double* cj = col + (j & 1) * NBK;
    cj[lane] = u;
    first_bad = (first_bad == 0 && j < nb && !(dj > 0.0)) ? k0 + j + 1 : first_bad;
    const double rcp = rcp_seeded(dj);
    const double nusq = -(u * u);                  // ready before the reciprocal is
    double dnext = 0.0;
    if (j + 1 < NBK) dnext = __shfl_sync(0xffffffffu, fma(nusq, rcp, a[(j + 1) % NBK]), (j + 1) % NBK);
    const double nt = -(u * rcp);
    __syncwarp();
#pragma unroll
    for (int c2 = (j + 1) / 2; c2 < NBK / 2; ++c2) {
      const double2 l = *reinterpret_cast<const double2*>(cj + 2 * c2);
      if (2 * c2 > j) a[2 * c2] = fma(nt, l.x, a[2 * c2]);
      a[2 * c2 + 1] = fma(nt, l.y, a[2 * c2 + 1]);
    }
    const double inv = rsqrt_seeded(dj);
    a[j] = lane == j ? dj * inv : u * inv;          // lanes above the diagonal carry unused values
    myinv = lane == j ? inv : myinv;
    dj = dnext;
  }
  *bad = first_bad;
  return myinv;
}

// z (one row per lane) <- z L11^-T for rows pre-scaled as z_c = x_c / L11[c][c], with
// Mt[q][c] = L11[c][q] / L11[c][c] in shared memory: one dependent fma per column.
__device__ __forceinline__ void panel_solve_rows(double (&z)[NBK], const double* Mt) {
#pragma unroll
  for (int c = 0; c < NBK; ++c) {
    const double nz = -z[c];
#pragma unroll
    for (int q2 = (c + 1) / 2; q2 < NBK / 2; ++q2) {
      const double2 l = *reinterpret_cast<const double2*>(Mt + c * NBK + 2 * q2);
      if (2 * q2 > c) z[2 * q2] = fma(nz, l.x, z[2 * q2]);
      z[2 * q2 + 1] = fma(nz, l.y, z[2 * q2 + 1]);
    }
  }
}

#ifdef CHOL_PROFILE
// tuning build (EDRGP_NVCC_EXTRA=-DCHOL_PROFILE): clock64 stamps of CTA 0's warp 0 per step
__device__ long long chol_prof[64][8];
#define CHOL_STAMP(i) do { if (blockIdx.x == 0 && tid == 0) chol_prof[kb & 63][i] = clock64(); } while (0)
#else
#define CHOL_STAMP(i) do { } while (0)
#endif

// Panel tiles live in shared memory swizzled, element (r, c) at r * 32 + (c ^ f(r)): a lane per row with
// a fixed column (the solves) and the DMMA fragment pattern (rows g, columns t) are both conflict free.
__device__ __forceinline__ int pswz(int r, int c) { return r * NBK + (c ^ (((r & 3) << 2) | ((r >> 2) & 3))); }

struct __align__(16) CholWarpSmem {
  double Mt[NBK * NBK];          // transposed diagonal factor, rows scaled by their diagonal entry
  double P[NBK * NBK];           // panel tile (swizzled): in A_ik, out L_ik
  double col[2 * NBK];
  double dinv[NBK];
};

// Step kb of the solve.  Row blocks kb+1 .. nblk-1 of A plus (rhs != NULL) the right-hand side as row
// block nblk; CTAs = trailing tiles (i >= j > kb) + one finisher that stores L_kk and the solved
// piece of the right-hand side (into cvec).
__global__ void __launch_bounds__(256) chol_step_kernel(double* __restrict__ A, int64_t ld, int m, int kb,
                                                        double* __restrict__ Lout, int64_t ldl,
                                                        double* __restrict__ rhs, double* __restrict__ cvec,
                                                        int* __restrict__ info) {
  __shared__ double D[NBK][NBK + 1];
  __shared__ CholWarpSmem ws[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (m + NBK - 1) / NBK;
  const int k0 = kb * NBK, nb = min(NBK, m - k0);
  const int nt = nblk - kb - 1;                       // trailing block columns
  const int nrow = nt + (rhs != nullptr ? 1 : 0);     // trailing block rows
  // blockIdx -> (ii, jj), ii >= jj, column by column; past the last tile: the finisher
  int jj = 0, rem = blockIdx.x;
  while (jj < nt && rem >= nrow - jj) { rem -= nrow - jj; ++jj; }
  const bool finisher = jj >= nt;
  const int ib = finisher ? nblk : kb + 1 + jj + rem; // row block of this CTA's tile (nblk = the rhs row)
  const int jb = kb + 1 + jj;
  const bool is_rhs = ib >= nblk;
  const bool diag = !finisher && ib == jb;
  const bool two = !finisher && !diag;                // a second panel tile (block row jb)
  // programmatic dependent launch: the next step's CTAs may be scheduled now (they wait below for this
  // grid to complete), which takes the launch latency out of the chain of dependent steps
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  CHOL_STAMP(0);
  // ---- stage the diagonal block and the panel tiles: every load in flight before the first store --------
  {
    double vd[4], vi[4], vj[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + 256 * it, r = e >> 5, c = e & 31;
      vd[it] = r == c ? 1.0 : 0.0;
      if (r < nb && c < nb && c <= r) vd[it] = A[(int64_t)(k0 + r) * ld + k0 + c];
      vi[it] = 0.0;
      if (c < nb) {
        if (is_rhs) { if (r == 0 && rhs != nullptr) vi[it] = rhs[k0 + c]; }
        else if (ib * NBK + r < m) vi[it] = A[(int64_t)(ib * NBK + r) * ld + k0 + c];
      }
      vj[it] = 0.0;
      if (two && c < nb && jb * NBK + r < m) vj[it] = A[(int64_t)(jb * NBK + r) * ld + k0 + c];
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + 256 * it, r = e >> 5, c = e & 31;
      D[r][c] = vd[it];
      ws[0].P[pswz(r, c)] = vi[it];
      if (two) ws[1].P[pswz(r, c)] = vj[it];
    }
  }
  // this thread's entries of the trailing tile in the DMMA accumulator layout: warp w owns rows
  // 8 (w / 2) + g and the two 8-column blocks at 16 (w % 2); lane = 4 g + t holds columns 2t, 2t + 1 of each
  const int g = lane >> 2, t4 = lane & 3;
  const int tr = 8 * (warp >> 1) + g, tcb = 16 * (warp & 1) + 2 * t4;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};               // (block 0: 2t, 2t + 1), (block 1: 2t, 2t + 1)
  if (!finisher) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int gc = jb * NBK + tcb + (cc >> 1) * 8 + (cc & 1);
      if (gc < m) {
        if (is_rhs) { if (tr == 0) acc[cc] = rhs[gc]; }
        else if (ib * NBK + tr < m) acc[cc] = A[(int64_t)(ib * NBK + tr) * ld + gc];
      }
    }
  }
  __syncthreads();
  CHOL_STAMP(1);
  // ---- warps 0 / 1: factor the diagonal block, solve one panel tile each -----------------------------
  if (warp < 2 && (warp == 0 || two)) {
    CholWarpSmem& w = ws[warp];
    double a[NBK];
#pragma unroll
    for (int c = 0; c < NBK; ++c) a[c] = D[lane][c];
    int bad;
    const double dinv = potf2_bcast(a, w.col, lane, nb, k0, &bad);
    if (finisher && lane == 0 && bad != 0) atomicCAS(info, 0, bad);
    CHOL_STAMP(2);
    w.dinv[lane] = dinv;
#pragma unroll
    for (int c = 0; c < NBK; ++c) w.Mt[c * NBK + lane] = a[c] * dinv;
    if (finisher) {
#pragma unroll
      for (int c = 0; c < NBK; ++c)
        if (c < nb && c <= lane && lane < nb) Lout[(int64_t)(k0 + lane) * ldl + k0 + c] = a[c];
    }
    __syncwarp();
    double z[NBK];
#pragma unroll
    for (int c = 0; c < NBK; ++c) z[c] = w.P[pswz(lane, c)] * w.dinv[c];
    panel_solve_rows(z, w.Mt);
#pragma unroll
    for (int c = 0; c < NBK; ++c) w.P[pswz(lane, c)] = z[c];
    CHOL_STAMP(3);
  }
  __syncthreads();
  CHOL_STAMP(4);
  if (finisher) {
    if (rhs != nullptr && tid < nb) cvec[k0 + tid] = ws[0].P[pswz(0, tid)];
    return;
  }
  // ---- panel output (first trailing column only) and the trailing update on the FP64 tensor pipe -------
  if (jj == 0 && !is_rhs) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + 256 * it, r = e >> 5, c = e & 31;
      if (c < nb && ib * NBK + r < m) Lout[(int64_t)(ib * NBK + r) * ldl + k0 + c] = ws[0].P[pswz(r, c)];
    }
  }
  const double* Pi = ws[0].P;
  const double* Pj = diag ? ws[0].P : ws[1].P;
  const int cb = 16 * (warp & 1);
#pragma unroll
  for (int q0 = 0; q0 < NBK; q0 += 4) {
    const double na = -Pi[pswz(tr, q0 + t4)];
    const double b0 = Pj[pswz(cb + g, q0 + t4)], b1 = Pj[pswz(cb + 8 + g, q0 + t4)];
    dmma(acc[0], acc[1], na, b0);
    dmma(acc[2], acc[3], na, b1);
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int lc = tcb + (cc >> 1) * 8 + (cc & 1), gc = jb * NBK + lc;
    if (gc >= m) continue;
    if (is_rhs) { if (tr == 0) rhs[gc] = acc[cc]; }
    else if (ib * NBK + tr < m && (!diag || lc <= tr)) A[(int64_t)(ib * NBK + tr) * ld + gc] = acc[cc];
  }
  CHOL_STAMP(5);
}

cudaError_t launch_trsm(const double* L, int m, int64_t ldl, double* B, int nrhs, int64_t ldb, int trans,
                        cudaStream_t st);

#ifdef CHOL_PROFILE
extern "C" int edrgp_debug_chol_prof(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, chol_prof, sizeof(chol_prof));
}
#endif

// A x = rhs by Cholesky: A (m, ld) lower triangle destroyed, L (m, ldl) receives the factor (lower
// triangle), x (m) the solution; rhs (m) is destroyed (may be NULL: factor only, x unused).
cudaError_t launch_posv(double* A, int m, int64_t ld, double* L, int64_t ldl, double* rhs, double* x, int* info,
                        cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int nblk = (m + NBK - 1) / NBK;
  for (int kb = 0; kb < nblk; ++kb) {
    const int nt = nblk - kb - 1, nrow = nt + (rhs != nullptr ? 1 : 0);
    const int tiles = nt * nrow - nt * (nt - 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(tiles + 1));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, chol_step_kernel, A, ld, m, kb, L, ldl, rhs, x, info); count_launch();
    if (e != cudaSuccess) return e;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (rhs != nullptr) return launch_trsm(L, m, ldl, x, 1, 1, 1, st);
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// Triangular solves with the lower factor L (m x m):  trans = 0: L X = B,  trans = 1: L^T X = B.
// B is m x nrhs row-major, overwritten with X.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) trsm_diag_kernel(const double* __restrict__ Lm, int64_t ldl, int k0, int nb,
                                                       double* __restrict__ B, int64_t ldb, int nrhs, int trans) {
  __shared__ double L[NBK][NBK + 1];
  const int i = threadIdx.x;
  for (int c = 0; c < nb; ++c) L[i][c] = (i < nb && c <= i) ? Lm[(int64_t)(k0 + i) * ldl + k0 + c] : 0.0;
  __syncwarp();
  const int col = blockIdx.x * 32 + i;
  if (col >= nrhs) return;
  double x[NBK];
#pragma unroll
  for (int r = 0; r < NBK; ++r) x[r] = r < nb ? B[(int64_t)(k0 + r) * ldb + col] : 0.0;
  if (!trans) {
#pragma unroll
    for (int r = 0; r < NBK; ++r) {
      if (r < nb) {
        double v = x[r];
#pragma unroll
        for (int q = 0; q < r; ++q) v = fma(-L[r][q], x[q], v);
        x[r] = v / L[r][r];
      }
    }
  } else {
#pragma unroll
    for (int r = NBK - 1; r >= 0; --r) {
      if (r < nb) {
        double v = x[r];
#pragma unroll
        for (int q = r + 1; q < NBK; ++q)
          if (q < nb) v = fma(-L[q][r], x[q], v);
        x[r] = v / L[r][r];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NBK; ++r)
    if (r < nb) B[(int64_t)(k0 + r) * ldb + col] = x[r];
}

// One right-hand side: the whole solve in ONE CTA (256 threads).  Per 32-block: warp 0 solves the
// diagonal block with the unknowns in registers (one per lane, shuffles), then all threads update
// the remaining entries.  x has stride incx.
__global__ void __launch_bounds__(256) trsv_kernel(const double* __restrict__ L, int64_t ldl, int m,
                                                   double* __restrict__ x, int64_t incx, int trans) {
  extern __shared__ double xs[];             // [m] the vector, then [32][33] diagonal block
  double* D = xs + ((m + 1) & ~1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < m; i += blockDim.x) xs[i] = x[i * incx];
  __syncthreads();
  const int nblk = (m + NBK - 1) / NBK;
  for (int bb = 0; bb < nblk; ++bb) {
    const int b = trans ? nblk - 1 - bb : bb;
    const int k0 = b * NBK, nb = min(NBK, m - k0);
    // diagonal block -> shared (identity padded): D[r][c] = L[k0+r][k0+c]
    for (int i = tid; i < NBK * NBK; i += blockDim.x) {
      const int r = i >> 5, c = i & 31;
      double v = r == c ? 1.0 : 0.0;
      if (r < nb && c <= r) v = L[(int64_t)(k0 + r) * ldl + k0 + c];
      D[r * (NBK + 1) + c] = v;
    }
    __syncthreads();
    if (warp == 0) {
      double xi = lane < nb ? xs[k0 + lane] : 0.0;
      const double inv = 1.0 / D[lane * (NBK + 1) + lane];
      if (!trans) {
#pragma unroll
        for (int j = 0; j < NBK; ++j) {
          const double xj = __shfl_sync(0xffffffffu, xi * inv, j);      // final x_j (lane j)
          if (lane == j) xi = xj;
          else if (lane > j) xi = fma(-D[lane * (NBK + 1) + j], xj, xi);
        }
      } else {
#pragma unroll
        for (int j = NBK - 1; j >= 0; --j) {
          const double xj = __shfl_sync(0xffffffffu, xi * inv, j);
          if (lane == j) xi = xj;
          else if (lane < j) xi = fma(-D[j * (NBK + 1) + lane], xj, xi);
        }
      }
      if (lane < nb) xs[k0 + lane] = xi;
    }
    __syncthreads();
    if (!trans) {
      // rows below: x_r -= sum_k L[r][k0+k] x_k
      for (int r = k0 + nb + tid; r < m; r += blockDim.x) {
        const double* lr = L + (int64_t)r * ldl + k0;
        double v = xs[r];
        double lv[NBK];
#pragma unroll
        for (int k = 0; k < NBK; ++k) lv[k] = k < nb ? lr[k] : 0.0;          // all loads in flight at once
        double v1 = 0.0;
#pragma unroll
        for (int k = 0; k < NBK; k += 2) {
          v = fma(-lv[k], k < nb ? xs[k0 + k] : 0.0, v);
          v1 = fma(-lv[k + 1], k + 1 < nb ? xs[k0 + k + 1] : 0.0, v1);
        }
        xs[r] = v + v1;
      }
    } else {
      // entries above: x_r -= sum_k L[k0+k][r] x_k   (coalesced over r)
      for (int r = tid; r < k0; r += blockDim.x) {
        double v = xs[r];
        double lv[NBK];
#pragma unroll
        for (int k = 0; k < NBK; ++k) lv[k] = k < nb ? L[(int64_t)(k0 + k) * ldl + r] : 0.0;
        double v1 = 0.0;
#pragma unroll
        for (int k = 0; k < NBK; k += 2) {
          v = fma(-lv[k], k < nb ? xs[k0 + k] : 0.0, v);
          v1 = fma(-lv[k + 1], k + 1 < nb ? xs[k0 + k + 1] : 0.0, v1);
        }
        xs[r] = v + v1;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < m; i += blockDim.x) x[i * incx] = xs[i];
}

// One right-hand side, m <= 704: the same solve with the 32 x 32 diagonal blocks INVERTED first (all
// blocks at once, a warp per block, a lane per column of the inverse: 32 dependent fma each), so that the
// sequential part of a block step is a 32 x 32 matrix-vector product (four partial sums: ~10 dependent
// FP64 operations) instead of 32 substitution steps of shuffle + multiply + fma (~3 300 cycles); the L
// entries of a block's update are fetched before the product, not after it.  Shared memory: the inverses
// (33 x 32 doubles per block) and the vector.
constexpr int TRSV_INV_MAX_BLOCKS = 22;
__global__ void __launch_bounds__(256) trsv_inv_kernel(const double* __restrict__ L, int64_t ldl, int m,
                                                       double* __restrict__ x, int64_t incx, int trans) {
  extern __shared__ __align__(16) double tsm[];
  const int nblk = (m + NBK - 1) / NBK;
  double* inv = tsm;                                   // [nblk][32][33]: inv[b][j * 33 + c] = (L_bb^-1)[c][j]
  double* xs = tsm + (size_t)nblk * NBK * (NBK + 1);   // [m]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < m; i += blockDim.x) xs[i] = x[i * incx];
  for (int b = warp; b < nblk; b += 8) {
    const int k0 = b * NBK, nb = min(NBK, m - k0);
    double* Mt = inv + (size_t)b * NBK * (NBK + 1);    // scratch for the scaled transposed block, then the inverse
    double a[NBK];
#pragma unroll
    for (int c = 0; c < NBK; ++c) {
      a[c] = c == lane ? 1.0 : 0.0;
      if (lane < nb && c <= lane) a[c] = L[(int64_t)(k0 + lane) * ldl + k0 + c];
    }
    double diag = 1.0;
#pragma unroll
    for (int c = 0; c < NBK; ++c) diag = c == lane ? a[c] : diag;
    const double dinv = 1.0 / diag;
#pragma unroll
    for (int c = 0; c < NBK; ++c) Mt[c * NBK + lane] = a[c] * dinv;
    __syncwarp();
    double z[NBK];
#pragma unroll
    for (int c = 0; c < NBK; ++c) z[c] = c == lane ? dinv : 0.0;
    panel_solve_rows(z, Mt);                            // z[c] = (L_bb^-1)[c][lane]
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NBK; ++c) Mt[lane * (NBK + 1) + c] = z[c];
  }
  __syncthreads();
  for (int bb = 0; bb < nblk; ++bb) {
    const int b = trans ? nblk - 1 - bb : bb;
    const int k0 = b * NBK, nb = min(NBK, m - k0);
    const double* ib = inv + (size_t)b * NBK * (NBK + 1);
    // this thread's first row / column of the update, fetched ahead of the block solve
    const int r0 = trans ? tid : k0 + nb + tid;
    const bool have = trans ? r0 < k0 : r0 < m;
    double lv[NBK];
    if (have) {
#pragma unroll
      for (int k = 0; k < NBK; ++k)
        lv[k] = k < nb ? (trans ? L[(int64_t)(k0 + k) * ldl + r0] : L[(int64_t)r0 * ldl + k0 + k]) : 0.0;
    }
    if (warp == 0) {
      double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < NBK; ++j) {
        // forward: (L^-1)[lane][j] = ib[j * 33 + lane];  backward: (L^-T)[lane][j] = (L^-1)[j][lane] = ib[lane * 33 + j]
        const double cij = trans ? ib[lane * (NBK + 1) + j] : ib[j * (NBK + 1) + lane];
        p[j & 3] = fma(cij, j < nb ? xs[k0 + j] : 0.0, p[j & 3]);
      }
      __syncwarp();
      if (lane < nb) xs[k0 + lane] = (p[0] + p[1]) + (p[2] + p[3]);
    }
    __syncthreads();
    for (int r = r0, pass = 0; trans ? r < k0 : r < m; r += blockDim.x, ++pass) {
      if (pass > 0) {
#pragma unroll
        for (int k = 0; k < NBK; ++k)
          lv[k] = k < nb ? (trans ? L[(int64_t)(k0 + k) * ldl + r] : L[(int64_t)r * ldl + k0 + k]) : 0.0;
      }
      double q[4] = {xs[r], 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < NBK; ++k) q[k & 3] = fma(-lv[k], k < nb ? xs[k0 + k] : 0.0, q[k & 3]);
      xs[r] = (q[0] + q[1]) + (q[2] + q[3]);
    }
    __syncthreads();
  }
  for (int i = tid; i < m; i += blockDim.x) x[i * incx] = xs[i];
}

cudaError_t launch_trsm(const double* L, int m, int64_t ldl, double* B, int nrhs, int64_t ldb, int trans,
                        cudaStream_t st) {
  const int nblk = (m + NBK - 1) / NBK;
  cudaError_t e = cudaSuccess;
  if (nrhs == 1 && nblk <= TRSV_INV_MAX_BLOCKS) {
    const size_t smem = ((size_t)nblk * NBK * (NBK + 1) + (size_t)m) * sizeof(double);
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(trsv_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    trsv_inv_kernel<<<1, 256, smem, st>>>(L, ldl, m, B, ldb, trans); count_launch();
    return cudaGetLastError();
  }
  if (nrhs == 1) {
    const size_t smem = ((size_t)((m + 1) & ~1) + NBK * (NBK + 1)) * sizeof(double);
    if (smem <= 200 * 1024) {
      if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(trsv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
      }
      trsv_kernel<<<1, 256, smem, st>>>(L, ldl, m, B, ldb, trans); count_launch();
      return cudaGetLastError();
    }
  }
  if (!trans) {
    for (int b = 0; b < nblk; ++b) {
      const int k0 = b * NBK, nb = min(NBK, m - k0);
      trsm_diag_kernel<<<(nrhs + 31) / 32, 32, 0, st>>>(L, ldl, k0, nb, B, ldb, nrhs, 0); count_launch();
      const int rest = m - k0 - nb;
      if (rest > 0) {   // B[rest] -= L[rest, blk] X_blk
        e = gemm_small(rest, nrhs, nb, -1.0, L + (int64_t)(k0 + nb) * ldl + k0, ldl, 1, B + (int64_t)k0 * ldb, ldb, 1,
                       1.0, B + (int64_t)(k0 + nb) * ldb, ldb, 0, st);
        if (e != cudaSuccess) return e;
      }
    }
  } else {
    for (int b = nblk - 1; b >= 0; --b) {
      const int k0 = b * NBK, nb = min(NBK, m - k0);
      trsm_diag_kernel<<<(nrhs + 31) / 32, 32, 0, st>>>(L, ldl, k0, nb, B, ldb, nrhs, 1); count_launch();
      if (k0 > 0) {     // B[0:k0] -= L[blk, 0:k0]^T X_blk
        e = gemm_small(k0, nrhs, nb, -1.0, L + (int64_t)k0 * ldl, 1, ldl, B + (int64_t)k0 * ldb, ldb, 1, 1.0, B, ldb, 0,
                       st);
        if (e != cudaSuccess) return e;
      }
    }
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// helpers for the VarDTC chain
// ---------------------------------------------------------------------------------------------
// Kmm from the kernel-entry matrix K(Z, Z): diagonal forced to sf2 (GPy zeroes r on the diagonal)
// plus jitter; the upper triangle is mirrored from the lower one.
__global__ void kmm_fix_kernel(double* __restrict__ K, int m, int64_t ld, double sf2, double jitter) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * m) return;
  const int r = (int)(idx / m), c = (int)(idx % m);
  if (r == c) K[r * ld + c] = sf2 + jitter;
  else if (c > r) K[r * ld + c] = K[c * ld + r];
}

// out = I + beta * 0.5 (T + T^T)
__global__ void make_B_kernel(const double* __restrict__ T, int m, double beta, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * m) return;
  const int r = (int)(idx / m), c = (int)(idx % m);
  const double v = beta * 0.5 * (T[(int64_t)r * m + c] + T[(int64_t)c * m + r]);
  out[idx] = v + (r == c ? 1.0 : 0.0);
}

__global__ void transpose_kernel(const double* __restrict__ in, int m, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (x < m && y0 + r < m) tile[r][threadIdx.x] = in[(int64_t)(y0 + r) * m + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (xo < m && yo0 + r < m) out[(int64_t)(yo0 + r) * m + xo] = tile[threadIdx.x][r];
}

// scalars[0] = trace(B) - m (= beta tr(Lm^-1 P Lm^-T)), scalars[1] = sum log diag(LB), scalars[2] = c^T c
__global__ void __launch_bounds__(256) solve_scalars_kernel(const double* __restrict__ Bmat, const double* __restrict__ LB,
                                                            const double* __restrict__ c, int m, double* __restrict__ scalars) {
  __shared__ double red[3][256];
  double a = 0.0, l = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) {
    a += Bmat[(int64_t)i * m + i] - 1.0;
    l += log(LB[(int64_t)i * m + i]);
    q = fma(c[i], c[i], q);
  }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = l; red[2][threadIdx.x] = q;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) scalars[threadIdx.x] = red[threadIdx.x][0];
}

__global__ void scale_copy_kernel(const double* __restrict__ in, double s, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * s;
}

// The VarDTC solve chain from the reduced statistics:
//   Lm = chol(Kmm);  A = beta Lm^-1 P Lm^-T;  LB = chol(I + A);
//   c = LB^-1 Lm^-1 (beta b);  alpha = Lm^-T LB^-T c.
// workspace: 2 m^2 doubles (T1, T2).  Kmm is overwritten by Lm (lower), Bmat by LB (lower).
cudaError_t launch_solve(double* Kmm, const double* P, const double* b, int m, double beta, double* Bmat,
                         double* alpha, double* cvec, double* scalars, int* info, double* workspace,
                         cudaStream_t st) {
  double* T1 = workspace;
  double* T2 = workspace + (size_t)m * m;
  const unsigned nb2 = (unsigned)(((int64_t)m * m + 255) / 256);
  cudaError_t e;
  if ((e = launch_potrf(Kmm, m, m, info, st)) != cudaSuccess) return e;
  // T1 = Lm^-1 P ; T2 = T1^T = P Lm^-T ; T2 <- Lm^-1 T2 = Lm^-1 P Lm^-T
  if ((e = cudaMemcpyAsync(T1, P, (size_t)m * m * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  if ((e = launch_trsm(Kmm, m, m, T1, m, m, 0, st)) != cudaSuccess) return e;
  dim3 tg((m + 31) / 32, (m + 31) / 32), tb(32, 8);
  transpose_kernel<<<tg, tb, 0, st>>>(T1, m, T2); count_launch();
  if ((e = launch_trsm(Kmm, m, m, T2, m, m, 0, st)) != cudaSuccess) return e;
  make_B_kernel<<<nb2, 256, 0, st>>>(T2, m, beta, Bmat); count_launch();
  // keep I + A for the trace before factorising in place: T1 <- B
  if ((e = cudaMemcpyAsync(T1, Bmat, (size_t)m * m * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  if ((e = launch_potrf(Bmat, m, m, info + 1, st)) != cudaSuccess) return e;
  scale_copy_kernel<<<(m + 255) / 256, 256, 0, st>>>(b, beta, m, cvec); count_launch();
  if ((e = launch_trsm(Kmm, m, m, cvec, 1, 1, 0, st)) != cudaSuccess) return e;
  if ((e = launch_trsm(Bmat, m, m, cvec, 1, 1, 0, st)) != cudaSuccess) return e;
  solve_scalars_kernel<<<1, 256, 0, st>>>(T1, Bmat, cvec, m, scalars); count_launch();
  if ((e = cudaMemcpyAsync(alpha, cvec, (size_t)m * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  if ((e = launch_trsm(Bmat, m, m, alpha, 1, 1, 1, st)) != cudaSuccess) return e;
  if ((e = launch_trsm(Kmm, m, m, alpha, 1, 1, 1, st)) != cudaSuccess) return e;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// The m x m part of the bound's hyper-parameter gradients (GPy VarDTC._compute_dL_dpsi / dL_dKmm,
// reached from edrgp/gp_model/base.py:69 on every L-BFGS evaluation):
//   E        = LB^-T (I + c c^T) LB^-1                      (DBi_plus_BiPBi)
//   dL_dpsi2 = beta/2 Lm^-T (I - E) Lm^-1
//   dL_dKmm  = Lm^-T (I - E/2 - B/2) Lm^-1,                 B = I + A as kept by the solve chain
//   sumAE    = sum (B - I) o E
// Outputs are the symmetrised forms the row pass and the Kuu part consume: Msym = (dL_dpsi2 + dL_dpsi2^T)/2 with
// leading dimension ldm, Dsym likewise (m x m, dense).  workspace: 3 m^2 doubles.
// ---------------------------------------------------------------------------------------------
__global__ void eye_plus_outer_kernel(const double* __restrict__ c, int m, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * m) return;
  const int r = (int)(idx / m), q = (int)(idx % m);
  out[idx] = fma(c[r], c[q], r == q ? 1.0 : 0.0);
}

// Et holds E^T: X2 = I - E, X3 = I - E/2 - B/2, partial sums of (B - I) o E per CTA (summed in order by the last step)
__global__ void __launch_bounds__(256) grad_small_mid_kernel(const double* __restrict__ Et, const double* __restrict__ B, int m,
                                                             double* __restrict__ X2, double* __restrict__ X3,
                                                             double* __restrict__ part) {
  __shared__ double red[256];
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double a = 0.0;
  if (idx < (int64_t)m * m) {
    const int r = (int)(idx / m), q = (int)(idx % m);
    const double e = Et[(int64_t)q * m + r], b = B[idx], id = r == q ? 1.0 : 0.0;
    X2[idx] = id - e;
    X3[idx] = id - 0.5 * e - 0.5 * b;
    a = (b - id) * e;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// out (ldo) = scale (T + T^T) / 2;  CTA 0 also folds the partial sums of the previous kernel into *sum
__global__ void __launch_bounds__(256) sym_scale_kernel(const double* __restrict__ T, int m, double scale, double* __restrict__ out,
                                                        int64_t ldo, const double* __restrict__ part, int nparts,
                                                        double* __restrict__ sum) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (int64_t)m * m) {
    const int r = (int)(idx / m), q = (int)(idx % m);
    out[(int64_t)r * ldo + q] = scale * 0.5 * (T[(int64_t)r * m + q] + T[(int64_t)q * m + r]);
  }
  if (sum != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < nparts; ++i) a += part[i];
    *sum = a;
  }
}

cudaError_t launch_vfe_grad_small(const double* LB, const double* Lm, const double* Bmat, const double* c, int m,
                                  double beta, double* Msym, int64_t ldm, double* Dsym, double* sumAE,
                                  double* workspace, cudaStream_t st) {
  double* W1 = workspace;
  double* W2 = workspace + (size_t)m * m;
  double* W3 = workspace + (size_t)2 * m * m;
  const unsigned nb2 = (unsigned)(((int64_t)m * m + 255) / 256);
  dim3 tg((m + 31) / 32, (m + 31) / 32), tb(32, 8);
  cudaError_t e;
  // L^-T X L^-1 for symmetric X: solve, transpose, solve (the result is left transposed; symmetric up to rounding)
  auto both_sides = [&](const double* L, double* X, double* Xt) -> cudaError_t {
    cudaError_t r = launch_trsm(L, m, m, X, m, m, 1, st);
    if (r != cudaSuccess) return r;
    transpose_kernel<<<tg, tb, 0, st>>>(X, m, Xt); count_launch();
    return launch_trsm(L, m, m, Xt, m, m, 1, st);
  };
  eye_plus_outer_kernel<<<nb2, 256, 0, st>>>(c, m, W1); count_launch();
  if ((e = both_sides(LB, W1, W2)) != cudaSuccess) return e;                       // W2 = E^T
  // the partial sums live at the head of Dsym until the last kernel has folded them
  grad_small_mid_kernel<<<nb2, 256, 0, st>>>(W2, Bmat, m, W1, W3, Dsym); count_launch();
  if ((e = both_sides(Lm, W1, W2)) != cudaSuccess) return e;                       // W2 = (2/beta dL_dpsi2)^T
  sym_scale_kernel<<<nb2, 256, 0, st>>>(W2, m, 0.5 * beta, Msym, ldm, Dsym, (int)nb2, sumAE); count_launch();
  if ((e = both_sides(Lm, W3, W1)) != cudaSuccess) return e;                       // W1 = dL_dKmm^T
  sym_scale_kernel<<<nb2, 256, 0, st>>>(W1, m, 1.0, Dsym, m, nullptr, 0, nullptr); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_kmm_fix(double* K, int m, int64_t ld, double sf2, double jitter, cudaStream_t st) {
  kmm_fix_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256, 0, st>>>(K, m, ld, sf2, jitter); count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigensolver (positive semi-definite input): one-sided cyclic Jacobi.
// ---------------------------------------------------------------------------------------------
// L lanes per column pair (32 / L pairs per warp), R rows per lane: d <= L R.  The solver is bound by
// the LATENCY of one round-robin step (shared-memory loads -> three dot products -> lane reduction ->
// rotation parameters -> column update -> block barrier), not by work: 64 columns give 32 independent
// pairs per step and ~600 dependent steps.  So the step is kept short: few rows per lane (more warps
// with shorter chains), the V columns are fetched before the parameters are derived, the pair schedule
// needs no division, and the rotation comes from two reciprocal square roots (no divide):
//     h = sqrt(diff^2 + 4 gamma^2), cos 2t = |diff| / h,  c = sqrt((1 + cos 2t) / 2),
//     s = sign(diff) gamma / (h c)            (|t| <= pi / 4, c^2 + s^2 = 1 to rounding)
// evaluated on operands scaled by a power of two so that the FP32 seeds stay in range.
// EDRGP_JACOBI_VARIANT (tuning aid, read once): unset = default (the d = 64 specialisation where it applies),
// 0 = the general kernel everywhere, 1 = fewer lanes per pair, 2 = a warp per pair, 3 = the d = 64 solver with the
// replay as a second kernel
static int jacobi_variant() {
  static const int v = [] { const char* e = getenv("EDRGP_JACOBI_VARIANT"); return e ? atoi(e) : -1; }();
  return v;
}

// Power-of-two factor that brings the largest diagonal entry of C into [1, 2): the solvers iterate on a scaled
// copy (rotations do not depend on the scale; eigenvalues are Rayleigh quotients against the caller's C), so that
// the squares and products of column norms they compare stay inside the double range for matrices whose entries are
// anywhere between ~1e-300 and ~1e+300.  The whole CTA calls this (one barrier); 1.0 for a zero or non-finite diagonal.
__device__ __forceinline__ double jacobi_input_scale(const double* __restrict__ C, int d, int tid, int nt) {
  __shared__ unsigned long long mxbits;
  if (tid == 0) mxbits = 0ull;
  __syncthreads();
  unsigned long long mine = 0ull;
  for (int i = tid; i < d; i += nt) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(C[(int64_t)i * d + i]));
    mine = b > mine ? b : mine;
  }
  if (mine) atomicMax(&mxbits, mine);
  __syncthreads();
  const int e = (int)((mxbits >> 52) & 0x7ffull);
  return (e == 0 || e == 0x7ff) ? 1.0 : __hiloint2double((2046 - e) << 20, 0);
}

// Eigenvalues and output for the vectors V (column i at V + i * ds, shared or global memory): lam is
// d doubles of shared scratch; the whole CTA calls this after a barrier.
__device__ __forceinline__ void jacobi_finish(const double* __restrict__ C, int d, const double* V, int ds, double* lam,
                                              double* __restrict__ evals, double* __restrict__ comps, int tid, int nt) {
  // eigenvalues: Rayleigh quotients v^T C v against the input (C is symmetric: read it by rows so that
  // the lanes of a warp touch consecutive addresses), one warp per vector, four independent chains
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int i0 = 0; i0 < d; i0 += nwarps) {
    const int i = i0 + warp;
    double acc = 0.0;
    if (i < d) {
      const double* v = V + i * ds;
      for (int r = lane; r < d; r += 32) {
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        int kk = 0;
        for (; kk + 3 < d; kk += 4) {
          c0 = fma(C[(int64_t)kk * d + r], v[kk], c0);
          c1 = fma(C[(int64_t)(kk + 1) * d + r], v[kk + 1], c1);
          c2 = fma(C[(int64_t)(kk + 2) * d + r], v[kk + 2], c2);
          c3 = fma(C[(int64_t)(kk + 3) * d + r], v[kk + 3], c3);
        }
        for (; kk < d; ++kk) c0 = fma(C[(int64_t)kk * d + r], v[kk], c0);
        acc = fma(v[r], (c0 + c1) + (c2 + c3), acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (i < d && lane == 0) lam[i] = acc;
  }
  __syncthreads();
  // sort descending by rank counting, write components as rows with a fixed sign (the first entry of
  // largest magnitude is made positive): a warp per vector
  for (int i = warp; i < d; i += nwarps) {
    const double li = lam[i];
    int rank = 0, bestj = d;
    double best = 0.0;
    for (int j = lane; j < d; j += 32) {
      const double lj = lam[j];
      rank += (lj > li) || (lj == li && j < i);
      const double v = V[i * ds + j];
      if (fabs(v) > fabs(best)) { best = v; bestj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rank += __shfl_xor_sync(0xffffffffu, rank, o);
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
      if (fabs(ob) > fabs(best) || (fabs(ob) == fabs(best) && oj < bestj)) { best = ob; bestj = oj; }
    }
    const double sgn = best < 0.0 ? -1.0 : 1.0;
    if (lane == 0) evals[rank] = li;
    for (int r = lane; r < d; r += 32) comps[(int64_t)rank * d + r] = sgn * V[i * ds + r];
  }
}


#ifdef JAC_PROFILE
// tuning build (EDRGP_NVCC_EXTRA=-DJAC_PROFILE): clock64 stamps of one thread over steps 4..7 of sweep 1
__device__ long long jac_prof[2][4][8];
#define JAC_STAMP(i) do { if (sweep == 1 && step >= 4 && step < 8 && (tid == 0 || tid == nt - 32)) \
    jac_prof[tid != 0][step - 4][i] = clock64(); } while (0)
extern "C" int edrgp_debug_jac_prof(long long* out) { return (int)cudaMemcpyFromSymbol(out, jac_prof, sizeof(jac_prof)); }
#else
#define JAC_STAMP(i) do { } while (0)
#endif

// LOGV: V is NOT carried here.  Its update never feeds back into the rotations, yet it is half of the
// shared-memory traffic and a fifth of the FP64 instructions of every step of this latency chain; the
// kernel only records (c, s) of every pair and step (rotlog[(sweep (dd - 1) + step) np + slot]; (1, 0)
// where nothing rotates) and the number of recorded sweeps (nlog), and jacobi_vectors_lean_kernel replays the
// log on the rows of V -- independent of one another -- across several SMs, then finishes (Rayleigh
// quotients, order, signs).
template <int L, int R, bool LOGV>
__global__ void __launch_bounds__(1024) jacobi_onesided_kernel(
    const double* __restrict__ C, int d, double* __restrict__ evals, double* __restrict__ comps, int max_sweeps,
    int* __restrict__ sweeps_out, double2* __restrict__ rotlog, int* __restrict__ nlog) {
  extern __shared__ double sh[];
  const int ds = d + 1 + ((d + 1) & 1);       // column stride: breaks the power-of-two bank pattern
  double* W = sh;                             // [d][ds] column-major: W[c * ds + r]
  double* V = LOGV ? sh : sh + (size_t)d * ds;
  __shared__ int rotated;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int dd = d + (d & 1), np = dd / 2;
  const double in_scale = jacobi_input_scale(C, d, tid, nt);
  for (int i = tid; i < d * d; i += nt) {
    const int c = i / d, r = i - c * d;
    W[c * ds + r] = C[(int64_t)r * d + c] * in_scale;
    if (!LOGV) V[c * ds + r] = r == c ? 1.0 : 0.0;
  }
  __syncthreads();
  const int k = tid / L, l = tid % L;         // pair slot and lane within the pair group
  const bool active = k < np;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < dd - 1; ++step) {
      int p = 0, q = d;
      JAC_STAMP(0);
      if (active) {                           // round-robin tournament; step, k < dd - 1: one conditional subtract
        int a0 = step + k, b0 = step + dd - 1 - k;
        if (a0 >= dd - 1) a0 -= dd - 1;
        if (b0 >= dd - 1) b0 -= dd - 1;
        if (k == 0) a0 = dd - 1;
        p = min(a0, b0); q = max(a0, b0);
      }
      const bool live = active && q < d;      // q == d: the bye of an odd dimension / idle slot
      double* Wp = W + p * ds + l;
      double* Wq = W + (live ? q : p) * ds + l;
      double* Vp = V + p * ds + l;
      double* Vq = V + (live ? q : p) * ds + l;
      constexpr bool HOIST = R <= 4 && !LOGV; // V columns fetched ahead of the rotation parameters (register budget)
      double wa[R], wb[R], va[HOIST ? R : 1], vb[HOIST ? R : 1];
      double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
      for (int e = 0; e < R; ++e) {
        const bool ok = live && l + L * e < d;
        wa[e] = ok ? Wp[L * e] : 0.0;
        wb[e] = ok ? Wq[L * e] : 0.0;
      }
      if (HOIST) {
#pragma unroll
        for (int e = 0; e < R; ++e) {
          const bool ok = live && l + L * e < d;
          va[HOIST ? e : 0] = ok ? Vp[L * e] : 0.0;
          vb[HOIST ? e : 0] = ok ? Vq[L * e] : 0.0;
        }
      }
      {
        // short dependent chains: NA partial sums per product, then a tree
        constexpr int NA = R >= 8 ? 4 : (R >= 2 ? 2 : 1);
        double pa[NA], pb[NA], pg[NA];
#pragma unroll
        for (int e = 0; e < NA; ++e) { pa[e] = wa[e] * wa[e]; pb[e] = wb[e] * wb[e]; pg[e] = wa[e] * wb[e]; }
#pragma unroll
        for (int e = NA; e < R; ++e) {
          pa[e % NA] = fma(wa[e], wa[e], pa[e % NA]); pb[e % NA] = fma(wb[e], wb[e], pb[e % NA]);
          pg[e % NA] = fma(wa[e], wb[e], pg[e % NA]);
        }
#pragma unroll
        for (int o = NA / 2; o > 0; o >>= 1) {
#pragma unroll
          for (int e = 0; e < o; ++e) { pa[e] += pa[e + o]; pb[e] += pb[e + o]; pg[e] += pg[e + o]; }
        }
        alpha = pa[0]; beta = pb[0]; gamma = pg[0];
      }
      JAC_STAMP(1);
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) {
        alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
        beta += __shfl_xor_sync(0xffffffffu, beta, o);
        gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
      }
      double c = 1.0, s = 0.0;
      JAC_STAMP(2);
      if (live && gamma * gamma > 1e-30 * alpha * beta && fabs(gamma) >= 1e-300) {
        const double sum = alpha + beta;
        const int ex = (__double2hiint(sum) >> 20) & 0x7ff;
        if (ex > 64 && ex < 1983) {                                 // 2^-959 < alpha + beta < 2^960
          const double scale = __hiloint2double((2046 - ex) << 20, 0);   // 2^(1023 - ex): sum * scale in [1, 2)
          const double dn = (beta - alpha) * scale, gn = 2.0 * gamma * scale;   // both in [-2, 2]
          const double ih = fast_rsqrt(fma(dn, dn, gn * gn));       // argument in [~1e-31, 8]
          const double x = fma(0.5 * fabs(dn), ih, 0.5);            // (1 + cos 2t) / 2 in [0.5, 1]
          const double r = fast_rsqrt(x);
          c = x * r;
          s = (dn >= 0.0 ? 0.5 : -0.5) * gn * ih * r;
        } else {                                                    // out of the scaled range: IEEE path
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = 1.0 / sqrt(fma(tt, tt, 1.0)); s = tt * c;
        }
        // a sweep whose rotations all stay below |cos| = 1e-9 leaves every pair orthogonal to ~1e-18 / gap
        // (quadratic convergence): it is the last one, no empty sweep is needed to find that out
        if (l == 0 && gamma * gamma > 1e-18 * alpha * beta) rotated = 1;
        JAC_STAMP(3);
#pragma unroll
        for (int e = 0; e < R; ++e) {
          if (l + L * e < d) {
            Wp[L * e] = c * wa[e] - s * wb[e];
            Wq[L * e] = s * wa[e] + c * wb[e];
            if (!LOGV) {
              const double x = HOIST ? va[HOIST ? e : 0] : Vp[L * e], y = HOIST ? vb[HOIST ? e : 0] : Vq[L * e];
              Vp[L * e] = c * x - s * y;
              Vq[L * e] = s * x + c * y;
            }
          }
        }
      }
      if (LOGV && active && l == 0) rotlog[((size_t)sweep * (dd - 1) + step) * np + k] = make_double2(c, s);
      JAC_STAMP(4);
      __syncthreads();
      JAC_STAMP(5);
    }
    if (!rotated) { ++sweep; break; }
    __syncthreads();
  }
  if (tid == 0 && sweeps_out) *sweeps_out = sweep;
  if (LOGV) {
    if (tid == 0) { nlog[0] = sweep; nlog[1] = 0; }      // sweeps recorded; the replay kernel's ticket
    return;
  }
  __syncthreads();
  jacobi_finish(C, d, V, ds, W, evals, comps, tid, nt);   // W's storage is free once every warp is past the barrier
}

#undef JAC_STAMP
// The one-sided solver with recorded rotations, specialised for d = 64: what bounds a
// step is the NUMBER of instructions each warp issues (228 in the general kernel, of which ~75 are the pair
// schedule, the `live` / row-bound predicates and address arithmetic: profiles/r01c_ncu_small_solvers.txt), so
// here the schedule of a whole sweep is a 4 KB shared-memory table built once, every pair is live, every lane
// owns exactly four rows and the log is written through a running pointer.  The arithmetic is the general
// kernel's, operation for operation (0.55 -> 0.51 ms on the bench's spectrum, profiles/r02_jacobi_variants.txt);
// the default for d = 64, EDRGP_JACOBI_VARIANT=0 selects the general kernel.
//
// ONE launch holds the solver (CTA 0) and the replay of its rotations on V (CTAs 1 .. 4, a warp per row of V): the
// replay used to be a second kernel that started when the solver had finished (0.08 ms of a 0.48 ms solve on the
// rank-replicated tail of the sweep); now it FOLLOWS the solver step by step.  The solver cannot afford a fence per
// step (a few hundred cycles in a step of ~1 700), so nothing is published: the log is filled with an all-ones
// pattern before the launch (one memset over log, V and control words), every entry is written with one 16-byte
// store, and the replay warps poll the entries of their next step past L1 until both words differ from the pattern.
// When the solver has stored its sweep count (ctrl[0]) and a warp has consumed that many sweeps, it is done; the last
// replay CTA (ticket) computes eigenvalues, order and signs as before.  Control words, all-ones = initial:
// [0] sweeps, [1] ticket, [3] time-out (cleared to 0 when a replay warp gave up after ~10 s).
constexpr int JD_THREADS = 512;
constexpr int JD_REPLAY_CTAS = 4;                       // 16 rows of V each
constexpr unsigned long long JD_EMPTY = 0xffffffffffffffffull;

__device__ __forceinline__ int ld_volatile_int(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void jacobi_d64_replay(const double* __restrict__ C, const double2* rotlog, int* ctrl,
                                                  double* Vt, double* __restrict__ evals, double* __restrict__ comps,
                                                  double* lam, int* last) {
  constexpr int D = 64, NP = 32, PER = 63;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = (blockIdx.x - 1) * 16 + warp;
  unsigned long long flip = 0;                          // bit st: the slot's first column is the larger index at step st
  for (int st = 0; st < PER; ++st) {
    int a0 = st + lane, b0 = st + PER - lane;
    if (a0 >= PER) a0 -= PER;
    if (b0 >= PER) b0 -= PER;
    if (lane == 0) a0 = PER;
    if (a0 > b0) flip |= 1ull << st;
  }
  const int a_init = lane == 0 ? PER : lane, b_init = lane == 0 ? 0 : PER - lane;
  double va = a_init == row ? 1.0 : 0.0, vb = b_init == row ? 1.0 : 0.0;
  const bool is0 = lane == 0, isl = lane == NP - 1;
  const double2* lp = rotlog + lane;
  int st = 0, g = 0, polls = 0;
  bool timed_out = false;
  const long long t_start = clock64();
  for (;;) {
    double c, sg;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(sg) : "l"(lp) : "memory");
    const bool ok = (unsigned long long)__double_as_longlong(c) != JD_EMPTY &&
                    (unsigned long long)__double_as_longlong(sg) != JD_EMPTY;
    if (!__all_sync(0xffffffffu, ok)) {
      // not written yet -- or the solver is done: its sweep count says how many steps there are
      int done = lane == 0 ? ld_volatile_int(ctrl) : 0;
      done = __shfl_sync(0xffffffffu, done, 0);
      if (done != -1 && g >= done * PER) break;
      if ((++polls & 1023) == 0 && clock64() - t_start > 20000000000ll) { timed_out = true; break; }   // ~10 s
      continue;
    }
    lp += NP;
    ++g;
    const unsigned fb = (unsigned)((flip >> st) & 1ull) << 31;
    sg = __hiloint2double(__double2hiint(sg) ^ (int)fb, __double2loint(sg));
    const double na = c * va - sg * vb, nb = sg * va + c * vb;
    const double dn = __shfl_down_sync(0xffffffffu, na, 1), up = __shfl_up_sync(0xffffffffu, nb, 1);
    va = is0 ? na : (isl ? nb : dn);
    vb = is0 ? dn : up;
    st = st + 1 == PER ? 0 : st + 1;
  }
  // a whole number of sweeps: st == 0, the slots hold their initial columns again
  Vt[(size_t)a_init * D + row] = va;
  Vt[(size_t)b_init * D + row] = vb;
  if (timed_out && lane == 0) ctrl[3] = 0;
  __threadfence();
  __syncthreads();
  if (tid == 0) *last = atomicAdd(reinterpret_cast<unsigned int*>(ctrl + 1), 1u) == (unsigned)(JD_REPLAY_CTAS - 2);
  __syncthreads();
  if (!*last) return;
  __threadfence();
  jacobi_finish(C, D, Vt, D, lam, evals, comps, tid, JD_THREADS);
  if (ld_volatile_int(ctrl + 3) != -1 && tid < D) evals[tid] = __longlong_as_double(0x7ff8000000000000ll);
}

__global__ void __launch_bounds__(JD_THREADS, 1) jacobi_d64_kernel(const double* __restrict__ C, int max_sweeps,
                                                                  int* __restrict__ sweeps_out, double2* __restrict__ rotlog,
                                                                  int* ctrl, double* Vt, double* __restrict__ evals,
                                                                  double* __restrict__ comps) {
  constexpr int D = 64, DS = 66, NP = 32, PER = 63, L = 16, R = 4;
  extern __shared__ double sh[];
  double* W = sh;                             // [D][DS] column-major
  __shared__ unsigned short sched[PER * NP];  // p | q << 8 of slot k at step s
  __shared__ int rotated;
  __shared__ int last;
  const int tid = threadIdx.x;
  if (blockIdx.x > 0) {
    jacobi_d64_replay(C, rotlog, ctrl, Vt, evals, comps, sh, &last);
    return;
  }
  const double in_scale = jacobi_input_scale(C, D, tid, JD_THREADS);
  for (int i = tid; i < D * D; i += JD_THREADS) {
    const int c = i >> 6, r = i & 63;
    W[c * DS + r] = C[(int64_t)r * D + c] * in_scale;
  }
  for (int i = tid; i < PER * NP; i += JD_THREADS) {
    const int step = i >> 5, k = i & 31;
    int a0 = step + k, b0 = step + PER - k;
    if (a0 >= PER) a0 -= PER;
    if (b0 >= PER) b0 -= PER;
    if (k == 0) a0 = PER;
    sched[i] = (unsigned short)(min(a0, b0) | (max(a0, b0) << 8));
  }
  __syncthreads();
  const int k = tid >> 4, l = tid & 15;
  double2* rl = rotlog + k;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < PER; ++step) {
      const unsigned pq = sched[step * NP + k];
      double* Wp = W + (pq & 0xffu) * DS + l;
      double* Wq = W + (pq >> 8) * DS + l;
      double wa[R], wb[R];
#pragma unroll
      for (int e = 0; e < R; ++e) { wa[e] = Wp[L * e]; wb[e] = Wq[L * e]; }
      double pa0 = wa[0] * wa[0], pb0 = wb[0] * wb[0], pg0 = wa[0] * wb[0];
      double pa1 = wa[1] * wa[1], pb1 = wb[1] * wb[1], pg1 = wa[1] * wb[1];
      pa0 = fma(wa[2], wa[2], pa0); pb0 = fma(wb[2], wb[2], pb0); pg0 = fma(wa[2], wb[2], pg0);
      pa1 = fma(wa[3], wa[3], pa1); pb1 = fma(wb[3], wb[3], pb1); pg1 = fma(wa[3], wb[3], pg1);
      double alpha = pa0 + pa1, beta = pb0 + pb1, gamma = pg0 + pg1;
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) {
        alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
        beta += __shfl_xor_sync(0xffffffffu, beta, o);
        gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
      }
      double c = 1.0, s = 0.0;
      if (gamma * gamma > 1e-30 * alpha * beta && fabs(gamma) >= 1e-300) {
        const double sum = alpha + beta;
        const int ex = (__double2hiint(sum) >> 20) & 0x7ff;
        if (ex > 64 && ex < 1983) {
          const double scale = __hiloint2double((2046 - ex) << 20, 0);
          const double dn = (beta - alpha) * scale, gn = 2.0 * gamma * scale;
          const double ih = fast_rsqrt(fma(dn, dn, gn * gn));
          const double x = fma(0.5 * fabs(dn), ih, 0.5);
          const double r = fast_rsqrt(x);
          c = x * r;
          s = (dn >= 0.0 ? 0.5 : -0.5) * gn * ih * r;
        } else {
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = 1.0 / sqrt(fma(tt, tt, 1.0)); s = tt * c;
        }
        if (l == 0 && gamma * gamma > 1e-18 * alpha * beta) rotated = 1;
#pragma unroll
        for (int e = 0; e < R; ++e) {
          Wp[L * e] = c * wa[e] - s * wb[e];
          Wq[L * e] = s * wa[e] + c * wb[e];
        }
      }
      if (l == 0) {
        // (a non-finite input can make the parameters NaN: the log must never hold the replay's "empty" pattern)
        const bool fin = fabs(c) <= 1.0 && fabs(s) <= 1.0;
        *rl = make_double2(fin ? c : 1.0, fin ? s : 0.0);
      }
      rl += NP;
      __syncthreads();
    }
    if (!rotated) { ++sweep; break; }
    __syncthreads();
  }
  if (tid == 0) {
    if (sweeps_out) *sweeps_out = sweep;
    asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(ctrl), "r"(sweep) : "memory");
  }
}

// Replays the rotation log of the solver on V = I.  A warp per ROW of V (rows never mix), lane = pair slot
// of the round-robin schedule holding that row's entries in the slot's two columns; between steps the columns
// move one slot along the tournament ring (two shuffles).  The last CTA to finish (ticket) computes the
// eigenvalues and writes the output.  d <= 64 (at most 32 slots).
// Per-step overhead is what this kernel is made of (4 FP64 operations, 4 shuffles and 1 load per step), so the
// orientation of a slot at every step of a sweep is a 63-bit mask computed once, the log is read through a
// running pointer sixteen steps ahead without bounds tests (the workspace carries 32 steps of slack) and only the
// tail batch is guarded: 26 instructions per step and warp (the first version, with the schedule recomputed
// and every access guarded, executed 64: profiles/r01c_ncu_small_solvers.txt; 0.11 -> 0.08 ms at d = 64,
// profiles/r02_jacobi_variants.txt).  Index logic emulated lane by lane on the CPU in tests/test_replay_logic.py.
constexpr int JV_WARPS = 8;
constexpr int JV_AHEAD = 16;
// V0 (may be null = identity; may alias Vt: a warp reads its own row before it writes it): the vectors the log
// continues from; finish = 0 leaves Vt as the result (an intermediate stage of the two-stage d = 64 solver).
__global__ void __launch_bounds__(JV_WARPS * 32) jacobi_vectors_lean_kernel(
    const double* __restrict__ C, int d, const double2* __restrict__ rotlog, int* ctrl, double* Vt,
    double* __restrict__ evals, double* __restrict__ comps, const double* V0, int finish) {
  __shared__ double lam[64];
  __shared__ int last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dd = d + (d & 1), np = dd / 2, per = dd - 1;
  const int row = blockIdx.x * JV_WARPS + warp;
  const int nsteps = ctrl[0] * per;
  if (row < d) {
    const bool act = lane < np;
    unsigned long long flip = 0;                    // bit st: the slot's first column is the larger index at step st
    for (int st = 0; st < per; ++st) {
      int a0 = st + lane, b0 = st + per - lane;
      if (a0 >= per) a0 -= per;
      if (b0 >= per) b0 -= per;
      if (lane == 0) a0 = per;
      if (act && a0 > b0) flip |= 1ull << st;
    }
    double va, vb;
    {
      const int a0 = lane == 0 ? per : lane, b0 = lane == 0 ? 0 : per - lane;
      if (V0 == nullptr) {
        va = (act && a0 == row) ? 1.0 : 0.0;
        vb = (act && b0 == row) ? 1.0 : 0.0;
      } else {
        va = (act && a0 < d) ? V0[(size_t)a0 * d + row] : 0.0;
        vb = (act && b0 < d) ? V0[(size_t)b0 * d + row] : 0.0;
      }
    }
    const bool is0 = lane == 0, isl = np > 1 && lane == np - 1, ring = np > 1;
    const double2* lp = rotlog + (act ? lane : 0);
    const size_t batch = (size_t)JV_AHEAD * np;
    double2 q[JV_AHEAD];
#pragma unroll
    for (int i = 0; i < JV_AHEAD; ++i) q[i] = lp[(size_t)i * np];
    lp += batch;
    int st = 0, g = 0;
#define JV_STEP(cs)                                                                                          \
    {                                                                                                        \
      const double c = (cs).x;                                                                               \
      const unsigned fb = (unsigned)((flip >> st) & 1ull) << 31;                                             \
      const double sg = __hiloint2double(__double2hiint((cs).y) ^ (int)fb, __double2loint((cs).y));          \
      const double na = c * va - sg * vb, nb = sg * va + c * vb;                                             \
      const double dn = __shfl_down_sync(0xffffffffu, na, 1), up = __shfl_up_sync(0xffffffffu, nb, 1);       \
      va = (is0 || !ring) ? na : (isl ? nb : dn);                                                            \
      vb = !ring ? nb : (is0 ? dn : up);                                                                     \
      st = st + 1 == per ? 0 : st + 1;                                                                       \
    }
    for (; g + JV_AHEAD <= nsteps; g += JV_AHEAD) {
#pragma unroll
      for (int i = 0; i < JV_AHEAD; ++i) {
        const double2 cs = q[i];
        q[i] = lp[(size_t)i * np];                  // 16 .. 32 steps ahead: inside the log or its slack
        JV_STEP(cs)
      }
      lp += batch;
    }
#pragma unroll
    for (int i = 0; i < JV_AHEAD; ++i) {
      if (g + i < nsteps) JV_STEP(q[i])             // uniform across the warp
    }
#undef JV_STEP
    if (act) {                                      // st == 0 again: the slots hold their initial columns
      const int a0 = lane == 0 ? per : lane, b0 = lane == 0 ? 0 : per - lane;
      if (a0 < d) Vt[(size_t)a0 * d + row] = va;
      if (b0 < d) Vt[(size_t)b0 * d + row] = vb;
    }
  }
  __threadfence();
  __syncthreads();
  if (!finish) return;
  if (tid == 0) last = atomicAdd(reinterpret_cast<unsigned int*>(ctrl + 1), 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  jacobi_finish(C, d, Vt, d, lam, evals, comps, tid, JV_WARPS * 32);
}

static size_t jacobi_log_doubles(int d) {      // 60 sweeps + 32 steps of slack (read ahead, never used)
  const int dd = d + (d & 1);
  return ((size_t)60 * (dd - 1) + 2 * JV_AHEAD) * (dd / 2) * 2;
}

template <int L, int R>
static cudaError_t launch_jacobi_small(const double* C, int d, double* ws, double* evals, double* comps, int* sweeps,
                                       cudaStream_t st) {
  const int ds = d + 1 + ((d + 1) & 1);
  const int np = (d + 1) / 2;
  int threads = ((np * L + 31) / 32) * 32;
  if (threads < 128) threads = 128;            // at least 4 warps for the Rayleigh / output phase
  if (d <= 64) {                               // rotations logged, vectors replayed across SMs
    double2* rotlog = reinterpret_cast<double2*>(ws);
    double* Vt = ws + jacobi_log_doubles(d);
    int* ctrl = reinterpret_cast<int*>(Vt + (size_t)d * d);          // [0] sweeps logged, [1] ticket
    const size_t smem = (size_t)d * ds * sizeof(double);
    // (a two-sided solver on the matrix itself was tried here: same time per step, fewer sweeps on flat spectra,
    // but its absolute rotation test does not terminate cleanly on numerically rank-deficient matrices)
    // (the same solver on a cluster of 2 or 4 SMs -- every block owning a share of the pairs, full column copies
    // kept coherent through distributed shared memory, one cluster barrier per step -- was built and measured in
    // round 2: 0.88 / 0.84 ms against 0.48 ms on one SM; a cluster barrier costs ~1 600 cycles, as much as the
    // whole step it was meant to shorten: profiles/r02_jacobi_variants.txt)
    const int nrep = (d + JV_WARPS - 1) / JV_WARPS;
    // (a single-precision pre-solve in front of the FP64 solver -- float copy of the matrix, rotations renormalised in
    // FP64 and logged, the FP64 solver warm-started from W = C V1 so that it needs two sweeps instead of eight -- was
    // built and measured in round 2: 0.487 ms against 0.486 ms.  A float step is not shorter: the step is a chain of
    // ~10 dependent stages (load, products, four shuffle levels, parameters, store, barrier), not FP64 issue.)
    if (jacobi_variant() == 3 && d == 64) {
      // tuning aid: the same solver, the replay as a second kernel behind it (what the default was before the replay
      // followed the solver inside one launch)
      cudaError_t e = cudaMemsetAsync(ctrl, 0, 2 * sizeof(int), st);
      if (e != cudaSuccess) return e;
      jacobi_d64_kernel<<<1, JD_THREADS, smem, st>>>(C, 60, sweeps, rotlog, ctrl, Vt, evals, comps);
      count_launch();
      jacobi_vectors_lean_kernel<<<nrep, JV_WARPS * 32, 0, st>>>(C, d, rotlog, ctrl, Vt, evals, comps, nullptr, 1);
      count_launch();
      return cudaGetLastError();
    }
    if (jacobi_variant() < 0 && d == 64) {
      // solver and replay in one launch; log, V and the control words start as the all-ones "empty" pattern
      cudaError_t e = cudaMemsetAsync(rotlog, 0xff, (jacobi_log_doubles(d) + (size_t)d * d + 2) * sizeof(double), st);
      if (e != cudaSuccess) return e;
      jacobi_d64_kernel<<<1 + JD_REPLAY_CTAS, JD_THREADS, smem, st>>>(C, 60, sweeps, rotlog, ctrl, Vt, evals, comps);
      count_launch();
      return cudaGetLastError();
    }
    jacobi_onesided_kernel<L, R, true><<<1, threads, smem, st>>>(C, d, evals, comps, 60, sweeps, rotlog, ctrl);
    count_launch();
    jacobi_vectors_lean_kernel<<<nrep, JV_WARPS * 32, 0, st>>>(C, d, rotlog, ctrl, Vt, evals, comps, nullptr, 1);
    count_launch();
    return cudaGetLastError();
  }
  const size_t smem = (size_t)2 * d * ds * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_onesided_kernel<L, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  jacobi_onesided_kernel<L, R, false><<<1, threads, smem, st>>>(C, d, evals, comps, 60, sweeps, nullptr, nullptr);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// One-sided Jacobi for large d (> 116): the same algorithm with W = C V and V in global memory
// (column-major, L2-resident), a warp per column pair, so that many SMs work on the d/2 independent
// pairs of a step; all steps and sweeps run inside one persistent kernel (grid barrier per step) and
// nothing synchronises the stream.
// ---------------------------------------------------------------------------------------------
__global__ void eig_init_kernel(const double* __restrict__ C, int d, double* __restrict__ W, double* __restrict__ V) {
  const double in_scale = jacobi_input_scale(C, d, threadIdx.x, blockDim.x);     // every CTA derives the same factor
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)d * d) return;
  const int c = (int)(i / d), r = (int)(i - (int64_t)c * d);
  W[i] = C[(int64_t)r * d + c] * in_scale;
  V[i] = r == c ? 1.0 : 0.0;
}

// All sweeps in ONE persistent kernel: a warp per column pair, a grid-wide barrier (global counter, every
// CTA resident: the grid never exceeds the SM count) after every round-robin step instead of a kernel
// launch (~11 us each, 6 600 of them at d = 512).  W and V live in global memory (L2) and are accessed
// with ld/st.global.cg: another CTA rewrote them one step ago.  ctrl[0] = barrier counter,
// ctrl[1 + (sweep & 1)] = "a rotation happened in this sweep", ctrl[3] = sweeps done, ctrl[4] = barrier
// time-out (never expected; it keeps a lost CTA from hanging the device).
__device__ __forceinline__ bool grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& target) {
  __syncthreads();
  __shared__ int ok_s;
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(counter, 1u);
    int ok = 1;
    unsigned int seen;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (++spins > 200000000LL) { ok = 0; break; }
    } while ((int)(seen - target) < 0);
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}

// EPL > 0: d <= 32 EPL and a lane keeps its EPL entries of all four columns (w_p, w_q, v_p, v_q) in
// registers: ONE L2 round trip per step (all loads in flight before the first use) instead of two
// passes of short dependent batches.  EPL = 0: any d, looping.
template <int EPL>
__global__ void __launch_bounds__(256) eig_persistent_kernel(double* __restrict__ W, double* __restrict__ V, int d,
                                                             int max_sweeps, unsigned int* __restrict__ ctrl) {
  const int dd = d + (d & 1), np = dd / 2;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5), nw = gridDim.x * wpb;
  unsigned int target = 0;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    unsigned int* flag = ctrl + 1 + (sweep & 1);
    for (int step = 0; step < dd - 1; ++step) {
      for (int k = gw; k < np; k += nw) {
        int a0 = step + k, b0 = step + dd - 1 - k;
        if (a0 >= dd - 1) a0 -= dd - 1;
        if (b0 >= dd - 1) b0 -= dd - 1;
        if (k == 0) a0 = dd - 1;
        const int p = min(a0, b0), q = max(a0, b0);
        if (q >= d) continue;
        double* wa = W + (int64_t)p * d;
        double* wb = W + (int64_t)q * d;
        double* va = V + (int64_t)p * d;
        double* vb = V + (int64_t)q * d;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        if constexpr (EPL > 0) {
          constexpr int E = EPL > 0 ? EPL : 1;
          double xa[E], xb[E], ua[E], ub[E];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int r = lane + 32 * e;
            const bool ok = r < d;
            xa[e] = ok ? __ldcg(wa + r) : 0.0;
            xb[e] = ok ? __ldcg(wb + r) : 0.0;
            ua[e] = ok ? __ldcg(va + r) : 0.0;
            ub[e] = ok ? __ldcg(vb + r) : 0.0;
          }
#pragma unroll
          for (int e = 0; e < E; ++e) {
            alpha = fma(xa[e], xa[e], alpha); beta = fma(xb[e], xb[e], beta); gamma = fma(xa[e], xb[e], gamma);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
            beta += __shfl_xor_sync(0xffffffffu, beta, o);
            gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
          }
          if (!(gamma * gamma > 1e-30 * alpha * beta) || fabs(gamma) < 1e-300) continue;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = fast_rsqrt(fma(tt, tt, 1.0)), s = tt * c;
          if (lane == 0) *flag = 1u;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int r = lane + 32 * e;
            if (r < d) {
              __stcg(wa + r, c * xa[e] - s * xb[e]); __stcg(wb + r, s * xa[e] + c * xb[e]);
              __stcg(va + r, c * ua[e] - s * ub[e]); __stcg(vb + r, s * ua[e] + c * ub[e]);
            }
          }
        } else {
        for (int r = lane; r < d; r += 32) {
          const double x = __ldcg(wa + r), y = __ldcg(wb + r);
          alpha = fma(x, x, alpha); beta = fma(y, y, beta); gamma = fma(x, y, gamma);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
          beta += __shfl_xor_sync(0xffffffffu, beta, o);
          gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
        }
        if (!(gamma * gamma > 1e-30 * alpha * beta) || fabs(gamma) < 1e-300) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = fast_rsqrt(fma(tt, tt, 1.0)), s = tt * c;
        if (lane == 0) *flag = 1u;
        for (int r = lane; r < d; r += 32) {
          const double x = __ldcg(wa + r), y = __ldcg(wb + r);
          __stcg(wa + r, c * x - s * y); __stcg(wb + r, s * x + c * y);
          const double u = __ldcg(va + r), v = __ldcg(vb + r);
          __stcg(va + r, c * u - s * v); __stcg(vb + r, s * u + c * v);
        }
        }
      }
      if (!grid_barrier(ctrl, gridDim.x, target)) { if (threadIdx.x == 0) ctrl[4] = 1u; return; }
    }
    // every CTA reads this sweep's flag after the last barrier; the other flag is cleared for the next sweep
    unsigned int rot;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(rot) : "l"(flag) : "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[1 + ((sweep + 1) & 1)] = 0u;
    if (!grid_barrier(ctrl, gridDim.x, target)) { if (threadIdx.x == 0) ctrl[4] = 1u; return; }
    if (rot == 0u) break;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[3] = (unsigned int)sweep;
}

// lam[i] = v_i^T C v_i (one warp per vector)
__global__ void __launch_bounds__(256) eig_rayleigh_kernel(const double* __restrict__ C, const double* __restrict__ V,
                                                           int d, double* __restrict__ lam) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= d) return;
  const double* v = V + (int64_t)i * d;
  double acc = 0.0;
  for (int r = lane; r < d; r += 32) {
    double cv = 0.0;
    for (int k = 0; k < d; ++k) cv = fma(C[(int64_t)r * d + k], v[k], cv);
    acc = fma(v[r], cv, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) lam[i] = acc;
}

__global__ void eig_finish_kernel(const double* __restrict__ lam, const double* __restrict__ V, int d,
                                  double* __restrict__ evals, double* __restrict__ comps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const double li = lam[i];
  int rank = 0;
  for (int j = 0; j < d; ++j) {
    const double lj = lam[j];
    rank += (lj > li) || (lj == li && j < i);
  }
  evals[rank] = li;
  double best = 0.0;
  for (int r = 0; r < d; ++r) {
    const double v = V[(int64_t)i * d + r];
    if (fabs(v) > fabs(best)) best = v;
  }
  const double sgn = best < 0.0 ? -1.0 : 1.0;
  for (int r = 0; r < d; ++r) comps[(int64_t)rank * d + r] = sgn * V[(int64_t)i * d + r];
}

size_t eigh_workspace_doubles(int d) {
  if (d <= 64) return jacobi_log_doubles(d) + (size_t)d * d + 2;      // rotation log, V, control words
  return d <= 116 ? (size_t)d * d : (size_t)2 * d * d + d + 4;
}

static cudaError_t launch_eigh_large(const double* C, int d, double* ws, double* evals, double* comps, int* sweeps,
                                     cudaStream_t st) {
  double* W = ws;
  double* V = ws + (size_t)d * d;
  double* lam = V + (size_t)d * d;
  unsigned int* ctrl = reinterpret_cast<unsigned int*>(lam + d);     // 8 words (eigh_workspace_doubles reserves 4 doubles)
  const int dd = d + (d & 1), np = dd / 2;
  eig_init_kernel<<<(unsigned)(((int64_t)d * d + 255) / 256), 256, 0, st>>>(C, d, W, V); count_launch();
  cudaError_t e = cudaMemsetAsync(ctrl, 0, 8 * sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  int grid = (np + 7) / 8;                       // 8 warps (pairs) per CTA; every CTA must be resident: <= one per SM
  if (grid > sms) grid = sms;
  // The kernel synchronises its CTAs with its own spin barrier, which is only safe when every CTA is resident at
  // the same time: a COOPERATIVE launch makes the driver guarantee that (or fail the launch with
  // cudaErrorCooperativeLaunchTooLarge) even when another stream, an NCCL kernel or another process holds SMs.
  int max_sweeps = 60;
  void* args[] = {(void*)&W, (void*)&V, (void*)&d, (void*)&max_sweeps, (void*)&ctrl};
  const void* kern = d <= 128 ? (const void*)eig_persistent_kernel<4>
                   : d <= 256 ? (const void*)eig_persistent_kernel<8>
                   : d <= 512 ? (const void*)eig_persistent_kernel<16> : (const void*)eig_persistent_kernel<0>;
  if ((e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(256), args, 0, st)) != cudaSuccess) return e;
  count_launch();
  if (sweeps) {      // [0] sweeps done, [1] status: the barrier's time-out flag (0 = clean)
    e = cudaMemcpyAsync(sweeps, ctrl + 3, 2 * sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  eig_rayleigh_kernel<<<(d + 7) / 8, 256, 0, st>>>(C, V, d, lam); count_launch();
  eig_finish_kernel<<<(d + 127) / 128, 128, 0, st>>>(lam, V, d, evals, comps); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_eigh(double* A, int d, double* V, double* evals, double* comps, int* sweeps, cudaStream_t st) {
  if (d <= 116) {       // two d x (d + 2) matrices in shared memory
    // lanes per pair x rows per lane: threads = pairs x lanes (d = 64: 32 x 16 = 512; d = 116: 58 x 16 = 928)
    const int variant = jacobi_variant();
    if (d <= 16) return launch_jacobi_small<4, 4>(A, d, V, evals, comps, sweeps, st);
    if (d <= 32) return variant == 1 ? launch_jacobi_small<4, 8>(A, d, V, evals, comps, sweeps, st)
                                     : launch_jacobi_small<8, 4>(A, d, V, evals, comps, sweeps, st);
    if (d <= 64) {
      if (variant == 1) return launch_jacobi_small<8, 8>(A, d, V, evals, comps, sweeps, st);
      if (variant == 2) return launch_jacobi_small<32, 2>(A, d, V, evals, comps, sweeps, st);
      return launch_jacobi_small<16, 4>(A, d, V, evals, comps, sweeps, st);
    }
    return launch_jacobi_small<16, 8>(A, d, V, evals, comps, sweeps, st);
  }
  return launch_eigh_large(A, d, V, evals, comps, sweeps, st);
}

}  // namespace edrgp
