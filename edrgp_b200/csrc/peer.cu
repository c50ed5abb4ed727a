// Rank-to-rank exchange of the sweep's three small reductions over NVLink peer memory (no NCCL on the data path).
//
// The fixed-hyper-parameter sweep reduces three payloads over the ranks: the targets' moments table (4 doubles per
// rank), the inducing statistics {P, b, y^T y} (m^2 + m + 1 doubles, 2.1 MB at m = 512) and the gradients' Gram matrix
// C (d^2 doubles).  Each is microseconds of traffic, so what a library collective costs here is not bandwidth but the
// launch, the host round trip between the composite calls and a kernel of its own between producer and consumer.  With
// one process per GPU on an NVSwitch box every rank can read every other rank's HBM directly, so the reductions are
// folded into the kernels that CONSUME them:
//
//   * every rank owns an exchange buffer obtained with cudaMalloc and exported with cudaIpcGetMemHandle; the peers map
//     it (cudaIpcOpenMemHandle).  Layout (doubles): flags | table x 2 | stats x 2 | gram x 2 -- two copies of every
//     payload, used alternately (parity of that payload's epoch), so that a rank can never overwrite a copy a slower
//     peer is still reading: to write parity p again it must have passed a wait of the epoch in between, which the
//     peer only signals after its own reads of p have completed (stream order).
//   * PUSH for the table: the rank writes its row into every peer's table and then, after a system-scope fence, its
//     epoch into every peer's flag word.  peer_table_wait_kernel spins on its own flags (local memory), then hands the table to the normaliser.
//   * PULL for P, b, y^T y and C: the producer kernels write the rank's partial straight into its own exchange buffer;
//     a one-warp kernel then raises the rank's flag on every peer.  The consumer -- form_system_peer_kernel, which also
//     assembles S = Kuu + beta P and beta b, and reduce_gram_peer_kernel in front of the eigensolver -- waits for all
//     flags and sums the peers' partials in RANK ORDER while it reads them over NVLink: every rank ends with bit-identical
//     P, b, C (and therefore alpha and components), as with the library collective.
//   * every wait is bounded (~10 s of clock64): a rank that never arrives sets bit 8 of the sweep's flag word instead of
//     hanging the device; the host raises on it.
//
// Reference: there is none -- the reference is single-process (edrgp/base.py:435-466 runs one estimator on all rows);
// this is the n-sharding of SURVEY 8(e), and replaces the three torch.distributed.all_reduce calls of model._fixed_pass /
// gradient_gram when all ranks of the job could map each other's buffers.
#include "common.cuh"
#include "launch.h"
#include "peer.h"

namespace edrgp {

__host__ __device__ static inline size_t even2(size_t x) { return x + (x & 1); }

size_t peer_layout(int m, int d, int world, int64_t* off) {
  size_t o = 0;
  off[PEER_FLAGS] = (int64_t)o; o += (size_t)PEER_COLLECTIVES * PEER_MAX_WORLD / 2;      // int32 words
  off[PEER_TABLE] = (int64_t)o; o += 2 * even2((size_t)4 * world);
  off[PEER_STATS] = (int64_t)o; o += 2 * even2((size_t)m * m + m + 1);
  off[PEER_GRAM] = (int64_t)o; o += 2 * even2((size_t)d * d);
  return o;
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double* p) {      // past L1, system scope: the line lives in a peer's L2
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Threads 0 .. world-1 of the calling CTA wait until rank r's flag of `coll` has reached `epoch`; everybody leaves
// together.  Returns false (for all threads) when a flag did not arrive in time.
__device__ __forceinline__ bool peer_wait(const PeerCtx& c, int coll, int epoch) {
  __shared__ int late;
  if (threadIdx.x == 0) late = 0;
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int* f = reinterpret_cast<const int*>(c.base[c.rank] + c.off[PEER_FLAGS]) + coll * PEER_MAX_WORLD + threadIdx.x;
    const long long t0 = clock64();
    int spins = 0;
    while (ld_acquire_sys(f) - epoch < 0) {
      if ((++spins & 255) == 0 && clock64() - t0 > 20000000000ll) { late = 1; break; }
    }
  }
  __syncthreads();
  return late == 0;
}

// one warp: make this rank's earlier writes visible system-wide, then raise its flag of `coll` on every rank
__global__ void __launch_bounds__(32) peer_signal_kernel(const PeerCtx c, int coll, int epoch) {
  __threadfence_system();
  if ((int)threadIdx.x < c.world)
    st_release_sys(reinterpret_cast<int*>(c.base[threadIdx.x] + c.off[PEER_FLAGS]) + coll * PEER_MAX_WORLD + c.rank, epoch);
}

// one warp: this rank's row of the moments table into every rank's table (copy `epoch & 1`), then the flag
__global__ void __launch_bounds__(32) peer_push_table_kernel(const PeerCtx c, const double* __restrict__ row, int epoch) {
  const int r = threadIdx.x;
  if (r < c.world) {
    double* dst = c.base[r] + c.off[PEER_TABLE] + (size_t)(epoch & 1) * even2((size_t)4 * c.world) + 4 * c.rank;
    const double v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(v0), "d"(v1) : "memory");
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(dst + 2), "d"(v2), "d"(v3) : "memory");
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(c.base[r] + c.off[PEER_FLAGS]) + PEER_COLL_TABLE * PEER_MAX_WORLD + c.rank, epoch);
  }
}

// target_standardize_kernel's front end under peer exchange: wait for every rank's row, then copy the table into the
// workspace (the kernel proper, and the host's deferred checks, read it there).  One CTA.
__global__ void __launch_bounds__(64) peer_table_wait_kernel(const PeerCtx c, int epoch, double* __restrict__ table,
                                                            unsigned int* __restrict__ flag) {
  const bool ok = peer_wait(c, PEER_COLL_TABLE, epoch);
  if (!ok && threadIdx.x == 0) atomicOr(flag, PEER_TIMEOUT_BIT);
  const double* src = c.base[c.rank] + c.off[PEER_TABLE] + (size_t)(epoch & 1) * even2((size_t)4 * c.world);
  for (int i = threadIdx.x; i < 4 * c.world; i += blockDim.x) table[i] = ld_peer(src + i);
}

// S (lower triangle and diagonal, mirrored) = K(Z, Z) with GPy's exact diagonal + jitter + beta P, rhs = beta b, with
// P, b, y^T y summed over the ranks' partials as they are read from the peers (rank order); the sums are also stored
// in `stats` (the workspace region the bound and the host read).  One thread per lower-triangle entry, as
// form_system_kernel; every CTA waits for the flags itself (they are local and, after the first CTA, already there).
__global__ void __launch_bounds__(256) form_system_peer_kernel(const PeerCtx c, int epoch, double* __restrict__ S, int m,
                                                               int64_t lds, double sf2, double jitter, double beta,
                                                               double* __restrict__ stats, double* __restrict__ rhs,
                                                               unsigned int* __restrict__ flag) {
  const bool ok = peer_wait(c, PEER_COLL_STATS, epoch);
  if (!ok && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flag, PEER_TIMEOUT_BIT);
  const size_t part = c.off[PEER_STATS] + (size_t)(epoch & 1) * even2((size_t)m * m + m + 1);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m + 1) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += ld_peer(c.base[r] + part + (size_t)m * m + idx);
    stats[(size_t)m * m + idx] = s;
    if (idx < m) rhs[idx] = beta * s;
  }
  if (idx >= (int64_t)m * m) return;
  const int r0 = (int)(idx / m), c0 = (int)(idx % m);
  if (c0 > r0) return;
  double p = 0.0;
  {
    double v[PEER_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r)
      if (r < c.world) v[r] = ld_peer(c.base[r] + part + (size_t)r0 * m + c0);       // all loads in flight, then the sum
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r)
      if (r < c.world) p += v[r];
  }
  stats[(size_t)r0 * m + c0] = p;
  if (c0 < r0) stats[(size_t)c0 * m + r0] = p;
  const double k = r0 == c0 ? sf2 + jitter : S[r0 * lds + c0];
  const double v = fma(beta, p, k);
  S[r0 * lds + c0] = v;
  if (c0 < r0) S[c0 * lds + r0] = v;
}

// C = sum over ranks of the partial Gram matrices (rank order), in front of the eigensolver
__global__ void __launch_bounds__(256) reduce_gram_peer_kernel(const PeerCtx c, int epoch, int count, double* __restrict__ C,
                                                               unsigned int* __restrict__ flag) {
  const bool ok = peer_wait(c, PEER_COLL_GRAM, epoch);
  if (!ok && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(flag, PEER_TIMEOUT_BIT);
  const size_t part = c.off[PEER_GRAM] + (size_t)(epoch & 1) * even2((size_t)count);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  double s = 0.0;
  for (int r = 0; r < c.world; ++r) s += ld_peer(c.base[r] + part + idx);
  C[idx] = s;
}

double* peer_partial(const PeerCtx& c, int region, int epoch, size_t count) {
  return c.base[c.rank] + c.off[region] + (size_t)(epoch & 1) * even2(count);
}

cudaError_t launch_peer_signal(const PeerCtx& c, int coll, int epoch, cudaStream_t st) {
  peer_signal_kernel<<<1, 32, 0, st>>>(c, coll, epoch); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_peer_push_table(const PeerCtx& c, const double* row, int epoch, cudaStream_t st) {
  peer_push_table_kernel<<<1, 32, 0, st>>>(c, row, epoch); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_peer_table_wait(const PeerCtx& c, int epoch, double* table, unsigned int* flag, cudaStream_t st) {
  peer_table_wait_kernel<<<1, 64, 0, st>>>(c, epoch, table, flag); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_form_system_peer(const PeerCtx& c, int epoch, double* S, int m, int64_t lds, double sf2, double jitter,
                                    double beta, double* stats, double* rhs, unsigned int* flag, cudaStream_t st) {
  form_system_peer_kernel<<<(unsigned)(((int64_t)m * m + 255) / 256), 256, 0, st>>>(c, epoch, S, m, lds, sf2, jitter, beta,
                                                                                   stats, rhs, flag);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_reduce_gram_peer(const PeerCtx& c, int epoch, int count, double* C, unsigned int* flag, cudaStream_t st) {
  reduce_gram_peer_kernel<<<(count + 255) / 256, 256, 0, st>>>(c, epoch, count, C, flag); count_launch();
  return cudaGetLastError();
}

}  // namespace edrgp
