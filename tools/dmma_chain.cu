// How many independent accumulator chains does the FP64 tensor pipe need?  Register-only DMMA loops with A
// independent accumulators per warp and W warps per scheduler (CTA of 4 W warps, one CTA per SM): TFLOP/s per
// (A, W).  If a DMMA.8x8x4 can only follow the previous DMMA on the SAME accumulator after L cycles, throughput
// saturates at min(1 / 16 cycles, W A / L) per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_chain tools/dmma_chain.cu && tools/dmma_chain
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int A>
__global__ void k(double* out, const double* in, int iters) {
  double acc[A][2];
#pragma unroll
  for (int i = 0; i < A; ++i) acc[i][0] = acc[i][1] = 0.0;
  const double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 32 / A; ++r)
#pragma unroll
      for (int i = 0; i < A; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < A; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int A>
static void run(int warps_per_sched, int sms, double* out, const double* in) {
  const int threads = 128 * warps_per_sched, iters = 4000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<A><<<sms, threads>>>(out, in, 100);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<A><<<sms, threads>>>(out, in, iters);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = (double)sms * (threads / 32) * iters * 32.0 * 512.0;
  printf("chains/warp %2d  warps/scheduler %d  -> %6.2f TFLOP/s\n", A, warps_per_sched, flops / (ms * 1e-3) / 1e12);
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double *out, *in; CK(cudaMalloc(&out, 1 << 16)); CK(cudaMalloc(&in, 1 << 16)); CK(cudaMemset(in, 0, 1 << 16));
  for (int w : {1, 2, 4}) {
    run<1>(w, sms, out, in); run<2>(w, sms, out, in); run<4>(w, sms, out, in); run<8>(w, sms, out, in);
    run<16>(w, sms, out, in); run<32>(w, sms, out, in);
  }
  return 0;
}
