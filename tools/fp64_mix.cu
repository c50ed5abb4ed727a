// What does a dependent FP64 operation cost next to a warp that keeps the FP64 tensor pipe busy?  One CTA of eight
// warps per SM: warps 0-3 (one per scheduler) issue DMMA.8x8x4 back to back (eight accumulator chains) or idle,
// warps 4-7 run C independent DFMA chains each.  Reported: cycles per DFMA of a chain (clock64 in the DFMA warp) and
// the DMMA warp's share of the tensor peak.  The answer decides how the exponentials of the cross-covariance kernels
// have to be arranged (profiles/r02_kuf2_study.txt): if a lone chain advances one link per ~30 cycles but C chains
// advance C links in the same time, the epilogue needs instruction-level parallelism, not fewer instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix tools/fp64_mix.cu && tools/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dfma(double a, double b, double c) {
  double d;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}

template <int C>
__global__ void k(double* out, const double* in, int iters, int with_dmma, long long* cyc) {
  const int warp = threadIdx.x >> 5;
  const double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  if (warp < 4) {
    if (!with_dmma) return;
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
    return;
  }
  double x[C];
#pragma unroll
  for (int i = 0; i < C; ++i) x[i] = a + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < C; ++i) x[i] = dfma(x[i], b, a);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += x[i];
  if (s == 123.456) out[threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 128) cyc[0] = t1 - t0;
}

template <int C>
static void run(int with_dmma, int sms, double* out, const double* in, long long* cyc) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<C><<<sms, 256>>>(out, in, 50, with_dmma, cyc);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<C><<<sms, 256>>>(out, in, iters, with_dmma, cyc);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  const double per_link = (double)c / (iters * 16.0);            // cycles for one link of every chain (C DFMAs)
  const double dmma_tf = with_dmma ? (double)sms * 4 * iters * 32.0 * 512.0 / (ms * 1e-3) / 1e12 : 0.0;
  printf("chains %2d  dmma warp %s : %6.1f cycles per link (= %5.1f per DFMA), kernel %7.3f ms, DMMA %5.2f TFLOP/s if the DMMA warps set the time\n",
         C, with_dmma ? "busy" : "idle", per_link, per_link / C, ms, dmma_tf);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double *out, *in; long long* cyc;
  CK(cudaMalloc(&out, 1 << 16)); CK(cudaMalloc(&in, 1 << 16)); CK(cudaMemset(in, 0, 1 << 16)); CK(cudaMalloc(&cyc, 8));
  for (int w : {0, 1}) {
    run<1>(w, sms, out, in, cyc); run<2>(w, sms, out, in, cyc); run<4>(w, sms, out, in, cyc); run<8>(w, sms, out, in, cyc);
    run<16>(w, sms, out, in, cyc);
  }
  return 0;
}
