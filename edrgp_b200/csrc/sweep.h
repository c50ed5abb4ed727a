// Internal declarations of the fixed-hyper-parameter sweep composites (sweep.cu, entry points in capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edrgp {

// regions of the shared workspace (the public header mirrors these as EDRGP_FS_*)
enum FixedRegion {
  FS_PACK_K = 0,   // inducing pack with unit coefficients (cross-covariance, Kuu)
  FS_PACK_G,       // inducing pack carrying alpha * std(y) (gradient pass)
  FS_YT,           // standardised targets of this rank
  FS_STATS,        // P (m x m) | b (m) | y^T y          -- all-reduced in place
  FS_TABLE,        // [n_r, pivot_r, S1_r, S2_r] per rank -- all-reduced in place (each rank fills its own row)
  FS_S,            // Kuu + jitter I + beta P (m x ld), consumed by the factorisation
  FS_L,            // its Cholesky factor
  FS_RHS,          // beta b (consumed)
  FS_ALPHA,        // posterior weights
  FS_SCRATCH,      // split-K partials / Gram partials / eigensolver workspace, one user at a time
  FS_TAIL,         // [u32 non-finite flag, i32 Cholesky info] | N | mean(y) | std(y)
  FS_RESULT,       // evals (d) | components (d x d) | C (d x d) | copy of the tail   -- one read-back
  FS_NREGIONS
};

// stats_mode: 0 = FP64 DMMA statistics, 1 = exact-product INT8 slices (i8syrk.cu; needs room for one block's slices)
size_t fixed_layout(int64_t n, int d, int m, int64_t chunk_rows, int world, int sms, int64_t* off, int stats_mode);
cudaError_t launch_target_moments(const double* y, int64_t n, double* part, unsigned int* ticket, double* slot, int sms,
                                  cudaStream_t st);
cudaError_t launch_target_standardize(const double* table, int world, const double* y, int64_t n, double* yt,
                                      double* tail3, int normalize, unsigned int* flag, int sms, cudaStream_t st);
cudaError_t launch_form_system(double* S, int m, int64_t lds, double sf2, double jitter, double beta, const double* P,
                               int64_t ldp, const double* b, double* rhs, cudaStream_t st);

}  // namespace edrgp
