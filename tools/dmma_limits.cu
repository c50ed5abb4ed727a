// What keeps a DMMA mainloop below the register-only peak?  Variants of the probe:
//   0: same A/B registers for every DMMA (the peak probe)
//   1: 4 distinct A x 8 distinct B registers (GEMM-like operand pattern), no memory
//   2: (1) + operands re-loaded from shared memory every k-step (12 LDS.64 per 32 DMMA)
//   3: (2) + __syncthreads every 4 k-steps
//   4: (3) + cp.async stage refill (global -> shared) like the SYRK kernel
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_limits tools/dmma_limits.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int iters) {
  extern __shared__ double sm[];     // [4 stages][2][16][132]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  for (int i = tid; i < 4 * 2 * 16 * 132; i += 256) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double a[4], b[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = in[lane + i];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = in[32 + lane + j];
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 4) {
      // refill one stage: 2 x 16 x 64 16-byte pieces = 8 per thread
      double* base = sm + (size_t)((it + 3) & 3) * 2 * 16 * 132;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = tid + q * 256;
        const int op = idx >> 10, rr = (idx >> 6) & 15, c2 = idx & 63;
        cp16(base + op * 16 * 132 + rr * 132 + 2 * c2, in + ((size_t)blockIdx.x * 4096 + (size_t)(it & 1023) * 2048 + idx * 2));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const double* As = sm + (size_t)(it & 3) * 2 * 16 * 132;
    const double* Bs = As + 16 * 132;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (MODE >= 2) {
        const double* ar = As + (4 * ks + t) * 132 + 32 * wm + g;
        const double* br = Bs + (4 * ks + t) * 132 + 64 * wn + g;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = ar[8 * i];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = br[8 * j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) dmma(acc[i][j][0], acc[i][j][1], a[0], b[0]);
          else dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    if (MODE >= 4) asm volatile("cp.async.wait_group 2;" ::: "memory");
    if (MODE >= 3) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[tid] = s;
}

template <int MODE>
static void run(int sms, double* out, double* in, int iters, int ctas) {
  const size_t smem = 4 * 2 * 16 * 132 * 8;
  CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = sms * ctas;
  for (int i = 0; i < 2; ++i) k<MODE><<<grid, 256, smem>>>(out, in, iters);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    CK(cudaEventRecord(e0)); k<MODE><<<grid, 256, smem>>>(out, in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double flops = (double)grid * 8 * iters * 128.0 * 512.0;
  printf("mode %d  ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", MODE, ctas, best, flops / (best * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double *in, *out; CK(cudaMalloc(&in, (size_t)p.multiProcessorCount * 2 * 4096 * 8 + 1024 * 2048 * 8 + (1 << 20))); CK(cudaMalloc(&out, 4096));
  CK(cudaMemset(in, 0, (size_t)p.multiProcessorCount * 2 * 4096 * 8 + 1024 * 2048 * 8 + (1 << 20)));
  for (int ctas = 1; ctas <= 1; ++ctas) {
    run<0>(p.multiProcessorCount, out, in, 4000, ctas);
    run<1>(p.multiProcessorCount, out, in, 4000, ctas);
    run<2>(p.multiProcessorCount, out, in, 4000, ctas);
    run<3>(p.multiProcessorCount, out, in, 4000, ctas);
    run<4>(p.multiProcessorCount, out, in, 4000, ctas);
  }
  return 0;
}
