// Small dense FP64 linear algebra on one GPU for the m x m inducing system and the d x d EDR
// matrix: blocked Cholesky, blocked triangular solves, cyclic Jacobi eigensolver.
//
// These stand in for the LAPACK calls GPy makes inside VarDTC.inference (dpotrf via jitchol,
// dtrtrs, dpotri) and for np.linalg.svd in SVDTransformer.fit (edrgp/utils.py:140).  Sizes are
// m <= a few thousand and d <= a few hundred: ~m^3 flops against the ~n m^2 of the statistics
// pass, so the kernels are written for clarity and determinism (32 x 32 blocks, no atomics), not
// for the last TFLOP.  All matrices are row-major.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int NBK = 32;

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
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) potf2_kernel(double* __restrict__ A, int64_t ld, int k0, int nb, int* info) {
  __shared__ double s[NBK][NBK + 1];
  const int i = threadIdx.x;
  for (int c = 0; c < nb; ++c) s[i][c] = (i < nb) ? A[(int64_t)(k0 + i) * ld + k0 + c] : 0.0;
  __syncwarp();
  for (int j = 0; j < nb; ++j) {
    const double ajj = s[j][j];
    if (!(ajj > 0.0)) {
      if (i == 0 && atomicCAS(info, 0, k0 + j + 1) == 0) {}
      return;
    }
    const double ljj = sqrt(ajj);
    __syncwarp();
    if (i == j) s[j][j] = ljj;
    if (i > j && i < nb) s[i][j] /= ljj;
    __syncwarp();
    const double lij = s[i][j];
    for (int c = j + 1; c < nb; ++c)
      if (i >= c && i < nb) s[i][c] = fma(-lij, s[c][j], s[i][c]);
    __syncwarp();
  }
  if (i < nb)
    for (int c = 0; c <= i; ++c) A[(int64_t)(k0 + i) * ld + k0 + c] = s[i][c];
}

// rows below the diagonal block:  A21 <- A21 L11^-T   (one thread per row)
__global__ void __launch_bounds__(32) potrf_panel_kernel(double* __restrict__ A, int64_t ld, int k0, int nb, int m,
                                                         const int* info) {
  __shared__ double L[NBK][NBK + 1];
  __shared__ double X[NBK][NBK + 1];
  if (*info != 0) return;
  const int i = threadIdx.x;
  const int row = k0 + nb + blockIdx.x * NBK + i;
  for (int c = 0; c < nb; ++c) {
    L[i][c] = (i < nb && c <= i) ? A[(int64_t)(k0 + i) * ld + k0 + c] : 0.0;
  }
  // coalesced load of the 32 x nb block of rows (lane = column)
  for (int r = 0; r < NBK; ++r) {
    const int rr = k0 + nb + blockIdx.x * NBK + r;
    X[r][i] = (rr < m && i < nb) ? A[(int64_t)rr * ld + k0 + i] : 0.0;
  }
  __syncwarp();
  if (row < m) {
    for (int c = 0; c < nb; ++c) {
      double v = X[i][c];
      for (int q = 0; q < c; ++q) v = fma(-X[i][q], L[c][q], v);
      X[i][c] = v / L[c][c];
    }
  }
  __syncwarp();
  for (int r = 0; r < NBK; ++r) {
    const int rr = k0 + nb + blockIdx.x * NBK + r;
    if (rr < m && i < nb) A[(int64_t)rr * ld + k0 + i] = X[r][i];
  }
}

cudaError_t launch_potrf(double* A, int m, int64_t ld, int* info, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  for (int k0 = 0; k0 < m; k0 += NBK) {
    const int nb = min(NBK, m - k0);
    potf2_kernel<<<1, 32, 0, st>>>(A, ld, k0, nb, info); count_launch();
    const int rest = m - k0 - nb;
    if (rest > 0) {
      potrf_panel_kernel<<<(rest + NBK - 1) / NBK, 32, 0, st>>>(A, ld, k0, nb, m, info); count_launch();
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

cudaError_t launch_trsm(const double* L, int m, int64_t ldl, double* B, int nrhs, int64_t ldb, int trans,
                        cudaStream_t st) {
  const int nblk = (m + NBK - 1) / NBK;
  cudaError_t e = cudaSuccess;
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

cudaError_t launch_kmm_fix(double* K, int m, int64_t ld, double sf2, double jitter, cudaStream_t st) {
  kmm_fix_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256, 0, st>>>(K, m, ld, sf2, jitter); count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigensolver: cyclic two-sided Jacobi with round-robin pair ordering, one CTA.
// A (d x d, destroyed) and V (d x d) live in global memory (L2-resident at these sizes).
// Output: evals sorted descending, comps[k][:] = k-th eigenvector (rows), sign fixed so that the
// largest-magnitude entry of every row is positive.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) jacobi_eigh_kernel(double* __restrict__ A, int d, double* __restrict__ V,
                                                           double* __restrict__ evals, double* __restrict__ comps,
                                                           int max_sweeps, int* __restrict__ sweeps_out) {
  extern __shared__ double sh[];
  const int dd = d + (d & 1);          // players in the round-robin (one bye if d is odd)
  const int np = dd / 2;
  double* cs_c = sh;                   // [np]
  double* cs_s = sh + np;              // [np]
  int* pp = reinterpret_cast<int*>(sh + 2 * np);   // [np]
  int* qq = pp + np;                   // [np]
  __shared__ int rotated;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < d * d; i += nt) V[i] = ((i / d) == (i % d)) ? 1.0 : 0.0;
  __syncthreads();
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < dd - 1; ++step) {
      // round-robin pairing: player dd-1 fixed, the others rotate
      for (int k = tid; k < np; k += nt) {
        int a = (k == 0) ? dd - 1 : (step + k) % (dd - 1);
        int b = (step + dd - 1 - k) % (dd - 1);
        int p = min(a, b), q = max(a, b);
        double c = 1.0, s = 0.0;
        if (q < d) {
          const double apq = A[(int64_t)p * d + q];
          const double app = A[(int64_t)p * d + p], aqq = A[(int64_t)q * d + q];
          if (fabs(apq) > 1e-300 && fabs(apq) > 2.220446049250313e-16 * 1e-2 * sqrt(fabs(app * aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt);
            s = tt * c;
            rotated = 1;
          }
        } else {
          p = -1;
        }
        cs_c[k] = c; cs_s[k] = s; pp[k] = p; qq[k] = q;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int i = tid; i < np * d; i += nt) {
        const int k = i / d, j = i - k * d;
        const int p = pp[k], q = qq[k];
        const double c = cs_c[k], s = cs_s[k];
        if (p >= 0 && s != 0.0) {
          const double x = A[(int64_t)p * d + j], y = A[(int64_t)q * d + j];
          A[(int64_t)p * d + j] = c * x - s * y;
          A[(int64_t)q * d + j] = s * x + c * y;
        }
      }
      __syncthreads();
      // columns: A <- A J,  V <- V J
      for (int i = tid; i < np * d; i += nt) {
        const int k = i % np, r = i / np;
        const int p = pp[k], q = qq[k];
        const double c = cs_c[k], s = cs_s[k];
        if (p >= 0 && s != 0.0) {
          double x = A[(int64_t)r * d + p], y = A[(int64_t)r * d + q];
          A[(int64_t)r * d + p] = c * x - s * y;
          A[(int64_t)r * d + q] = s * x + c * y;
          x = V[(int64_t)r * d + p]; y = V[(int64_t)r * d + q];
          V[(int64_t)r * d + p] = c * x - s * y;
          V[(int64_t)r * d + q] = s * x + c * y;
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  if (tid == 0 && sweeps_out) *sweeps_out = sweep;
  // sort descending by rank counting, write components as rows with a fixed sign
  for (int i = tid; i < d; i += nt) {
    const double li = A[(int64_t)i * d + i];
    int rank = 0;
    for (int j = 0; j < d; ++j) {
      const double lj = A[(int64_t)j * d + j];
      rank += (lj > li) || (lj == li && j < i);
    }
    evals[rank] = li;
    double best = 0.0;
    for (int r = 0; r < d; ++r) {
      const double v = V[(int64_t)r * d + i];
      if (fabs(v) > fabs(best)) best = v;
    }
    const double sgn = best < 0.0 ? -1.0 : 1.0;
    for (int r = 0; r < d; ++r) comps[(int64_t)rank * d + r] = sgn * V[(int64_t)r * d + i];
  }
}

cudaError_t launch_eigh(double* A, int d, double* V, double* evals, double* comps, int* sweeps, cudaStream_t st) {
  const int dd = d + (d & 1), np = dd / 2;
  const size_t smem = (size_t)np * (2 * sizeof(double) + 2 * sizeof(int));
  jacobi_eigh_kernel<<<1, 1024, smem, st>>>(A, d, V, evals, comps, 60, sweeps); count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
