// FP64 tensor-pipe peak probe: register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) accumulator
// chains with no memory traffic.  bench.py times it with CUDA events on the same box and in the
// same process as the workload, so the roofline denominator is a live measurement
// (MEASURED_PEAKS.json carries no FP64 figure).  tools/fp64_peak.cu is the standalone version.
#include "common.cuh"
#include "launch.h"

namespace edrgp {

constexpr int PROBE_CHAINS = 16;

__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters) {
  double a = 1.0 + 1e-9 * (threadIdx.x & 31), b = 1.0 - 1e-9 * (threadIdx.x & 31);
  double c0[PROBE_CHAINS], c1[PROBE_CHAINS];
#pragma unroll
  for (int u = 0; u < PROBE_CHAINS; ++u) { c0[u] = 0.0; c1[u] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < PROBE_CHAINS; ++u) dmma(c0[u], c1[u], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < PROBE_CHAINS; ++u) s += c0[u] + c1[u];
  if (s == 123.456) out[threadIdx.x] = s;      // never true: keeps the chains alive
}

// launches the probe; *flops receives the FP64 flops it executes (FMA = 2)
cudaError_t launch_dmma_probe(double* scratch, int iters, int sms, double* flops, cudaStream_t st) {
  const int grid = sms * 2;
  dmma_probe_kernel<<<grid, 256, 0, st>>>(scratch, iters); count_launch();
  if (flops) *flops = (double)grid * 8.0 * (double)iters * PROBE_CHAINS * 512.0;
  return cudaGetLastError();
}

}  // namespace edrgp
