// Internal declarations of the host-row streamer (h2d.cu; entry points edrgp_h2d_* in capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace edrgp {
struct H2DTransfer;
H2DTransfer* h2d_open(const void* src, void* dst, int64_t rows, size_t row_bytes, size_t dst_pitch, int64_t block_rows,
                      int threads, int slots, cudaStream_t order_after, const void* side_src, void* side_dst,
                      size_t side_bytes, cudaError_t* err);
cudaError_t h2d_wait_side(H2DTransfer* t, cudaStream_t consumer);
cudaError_t h2d_wait(H2DTransfer* t, int64_t upto_row, int64_t ahead_rows, cudaStream_t consumer);
bool h2d_staged(const H2DTransfer* t);
void h2d_close(H2DTransfer* t);
}  // namespace edrgp
