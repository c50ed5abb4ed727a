// Internal declarations of the NVLink peer-memory exchange (peer.cu; entry points edrgp_peer_* in capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edrgp {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_COLLECTIVES = 4;                    // table, stats, gram (+ one spare row of flags)
enum PeerColl { PEER_COLL_TABLE = 0, PEER_COLL_STATS = 1, PEER_COLL_GRAM = 2 };
enum PeerRegion { PEER_FLAGS = 0, PEER_TABLE, PEER_STATS, PEER_GRAM, PEER_NREGIONS };
constexpr unsigned int PEER_TIMEOUT_BIT = 0x100u;      // in the sweep's non-finite flag word: a rank never arrived

// passed to the kernels by value
struct PeerCtx {
  double* base[PEER_MAX_WORLD];      // every rank's exchange buffer as mapped into THIS process (base[rank] = own)
  int64_t off[PEER_NREGIONS];        // region offsets in doubles (peer_layout)
  int rank, world;
};

size_t peer_layout(int m, int d, int world, int64_t* off);      // doubles
// this rank's copy of a payload for `epoch` (count doubles per copy)
double* peer_partial(const PeerCtx& c, int region, int epoch, size_t count);
cudaError_t launch_peer_signal(const PeerCtx& c, int coll, int epoch, cudaStream_t st);
cudaError_t launch_peer_push_table(const PeerCtx& c, const double* row, int epoch, cudaStream_t st);
cudaError_t launch_peer_table_wait(const PeerCtx& c, int epoch, double* table, unsigned int* flag, cudaStream_t st);
cudaError_t launch_form_system_peer(const PeerCtx& c, int epoch, double* S, int m, int64_t lds, double sf2, double jitter,
                                    double beta, double* stats, double* rhs, unsigned int* flag, cudaStream_t st);
cudaError_t launch_reduce_gram_peer(const PeerCtx& c, int epoch, int count, double* C, unsigned int* flag, cudaStream_t st);

}  // namespace edrgp
