// Microbenchmark: issue rate of tcgen05.mma.kind::tf32 (M = 128, K = 8, both operands in shared memory,
// 128-byte-swizzled K-major tiles) on B200, one CTA per SM.  Prints cycles per MMA for several N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int distinct, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * 32768 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32768;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // `distinct` different k-slices / images per operand, as the kernels use them (1: same operands every time)
      const int v = r % distinct;
      const uint32_t off = (uint32_t)((v & 3) * 32 + ((v >> 2) & 1) * 16384);
      mma(tmem, desc_sw128(a0 + off), desc_sw128(b0 + off), idesc, r > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  const size_t smem = 2 * 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int reps = 4000;
  for (int grid : {1, sms}) {
    for (int N : {64, 80, 128, 256}) {
      for (int distinct : {1, 8}) {
        long long cyc = 0;
        for (int it = 0; it < 2; ++it) {
          rate_kernel<<<grid, 128, smem>>>(N, reps, distinct, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        }
        printf("{\"ctas\": %d, \"N\": %d, \"distinct_operands\": %d, \"cycles_per_mma\": %.1f, \"floor_MN_over_256\": %.1f}\n",
               grid, N, distinct, (double)cyc / reps, 128.0 * N / 256.0);
      }
    }
  }
  return 0;
}
