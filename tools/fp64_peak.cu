// FP64 peak microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 figures only; the EDR pipeline is bound by the FP64
// pipes, so its roofline denominator has to be measured on the box.  This program times
//   (1) DFMA  : register-resident independent FMA chains on the FP64 vector pipe,
//   (2) DMMA  : mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) independent accumulator chains,
//   (3) MIX   : both interleaved, to see whether the two pipes add up,
// with CUDA events, and prints one JSON line.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int U>
__global__ void __launch_bounds__(256) k_dfma(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) c[u] = (double)u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = fma(c[u], a, b);
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) s += c[u];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int U>
__global__ void __launch_bounds__(256) k_dmma(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c0[U], c1[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { c0[u] = 0; c1[u] = 0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < U; ++u) dmma884(c0[u], c1[u], a, b);
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) s += c0[u] + c1[u];
  if (s == 123.456) out[threadIdx.x] = s;
}

// U DMMA + V DFMA per iteration
template <int U, int V>
__global__ void __launch_bounds__(256) k_mix(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c0[U], c1[U], f[V];
#pragma unroll
  for (int u = 0; u < U; ++u) { c0[u] = 0; c1[u] = 0; }
#pragma unroll
  for (int v = 0; v < V; ++v) f[v] = (double)v;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dmma884(c0[u], c1[u], a, b);
#pragma unroll
      for (int v = u * V / U; v < (u + 1) * V / U; ++v) f[v] = fma(f[v], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) s += c0[u] + c1[u];
#pragma unroll
  for (int v = 0; v < V; ++v) s += f[v];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  double *in, *out;
  CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, 4096));
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));

  const int ctas_per_sm[3] = {1, 2, 4};
  double best_dfma = 0, best_dmma = 0, best_mix_tot = 0, best_mix_dmma = 0, best_mix_dfma = 0;
  for (int ci = 0; ci < 3; ++ci) {
    int grid = sms * ctas_per_sm[ci];
    double threads = (double)grid * 256, warps = threads / 32;
    {
      double ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, in, iters); }, 5);
      double tf = threads * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
      fprintf(stderr, "dfma  ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", ctas_per_sm[ci], ms, tf);
      if (tf > best_dfma) best_dfma = tf;
    }
    {
      double ms = time_ms([&] { k_dmma<16><<<grid, 256>>>(out, in, iters); }, 5);
      double tf = warps * iters * 16.0 * 512.0 / (ms * 1e-3) / 1e12;
      fprintf(stderr, "dmma  ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", ctas_per_sm[ci], ms, tf);
      if (tf > best_dmma) best_dmma = tf;
    }
    {
      // 8 DMMA (8*256 FMA per warp) + 64 DFMA-warp-instr (64*32 FMA per warp): equal flop on each pipe
      double ms = time_ms([&] { k_mix<8, 64><<<grid, 256>>>(out, in, iters); }, 5);
      double tfm = warps * iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;
      double tff = threads * iters * 64.0 * 2.0 / (ms * 1e-3) / 1e12;
      fprintf(stderr, "mix   ctas/sm=%d  %.3f ms  dmma %.2f + dfma %.2f = %.2f TFLOP/s\n",
              ctas_per_sm[ci], ms, tfm, tff, tfm + tff);
      if (tfm + tff > best_mix_tot) { best_mix_tot = tfm + tff; best_mix_dmma = tfm; best_mix_dfma = tff; }
    }
    {
      // DMMA-heavy mix: 8 DMMA + 16 DFMA (the shape of a GEMM mainloop with a light epilogue)
      double ms = time_ms([&] { k_mix<8, 16><<<grid, 256>>>(out, in, iters); }, 5);
      double tfm = warps * iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;
      double tff = threads * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
      fprintf(stderr, "mix8:16 ctas/sm=%d  %.3f ms  dmma %.2f + dfma %.2f = %.2f TFLOP/s\n",
              ctas_per_sm[ci], ms, tfm, tff, tfm + tff);
    }
  }
  int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"fp64_dfma_tflops\": %.3f, "
         "\"fp64_dmma_tflops\": %.3f, \"fp64_mix_tflops\": %.3f, \"fp64_mix_dmma\": %.3f, \"fp64_mix_dfma\": %.3f}\n",
         p.name, sms, clk, best_dfma, best_dmma, best_mix_tot, best_mix_dmma, best_mix_dfma);
  return 0;
}
