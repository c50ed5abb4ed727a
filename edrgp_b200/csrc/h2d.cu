// Host rows -> device, block by block, overlapped with the kernels that consume them (edrgp_h2d_*).
//
// The reference's estimator receives host arrays (edrgp/gp_model/base.py:46-91: check_X_y returns a C-contiguous
// float64 ndarray); at the headline shape that is 2 GB per fit, which takes as long to cross PCIe as the whole
// sweep takes to compute.  The Python host mirror therefore streams the rows in blocks and runs the statistics
// pass on the blocks that have arrived.  What happens to a block depends on where the caller's array lives:
//
//   pinned / registered memory   one cudaMemcpyAsync per block straight from the caller's buffer;
//   ordinary (pageable) memory   a cudaMemcpyAsync from pageable memory is staged by the driver through its own
//                                bounce buffer, one chunk at a time and synchronously with the calling thread.  Here
//                                worker threads copy stripes of the block into a ring of pinned slots (memcpy at DRAM
//                                speed, all workers in parallel), and the worker that completes a block enqueues its
//                                DMA and records the block's event -- the host copy of block b + 1 overlaps the DMA of
//                                block b and the kernels on block b - 1.
//
// Consumers never see a block early: edrgp_h2d_wait blocks the calling host thread until the block has been
// ENQUEUED (an event that has not been recorded yet would read as complete) and then makes the consumer stream wait
// for its event.  Rows may be narrower on the host than on the device (odd d padded to even): blocks are packed on
// the host side and copied with a 2-D copy.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "h2d.h"

namespace edrgp {

void host_copy_streaming(void* dst, const void* src, size_t n);      // hostcopy.cpp: non-temporal stores

namespace {

struct PinnedSlot { void* ptr = nullptr; size_t bytes = 0; };

// pinned slots are expensive to create (cudaHostAlloc pins and maps): released transfers keep them for the next one
struct SlotPool {
  std::mutex mu;
  std::vector<PinnedSlot> free_slots;
  PinnedSlot take(size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < free_slots.size(); ++i) {
        if (free_slots[i].bytes >= bytes) {
          PinnedSlot s = free_slots[i];
          free_slots.erase(free_slots.begin() + (long)i);
          return s;
        }
      }
    }
    PinnedSlot s;
    if (cudaHostAlloc(&s.ptr, bytes, cudaHostAllocDefault) != cudaSuccess) { s.ptr = nullptr; return s; }
    s.bytes = bytes;
    return s;
  }
  void give(PinnedSlot s) {
    if (!s.ptr) return;
    std::lock_guard<std::mutex> lk(mu);
    if (free_slots.size() < 8) free_slots.push_back(s);
    else cudaFreeHost(s.ptr);
  }
};
SlotPool g_pool;

}  // namespace

struct H2DTransfer {
  const char* src = nullptr;       // host rows, row_bytes each, contiguous
  char* dst = nullptr;             // device rows, dst_pitch each
  int64_t rows = 0;
  size_t row_bytes = 0, dst_pitch = 0;
  int64_t block_rows = 0;
  int nblocks = 0;
  bool staged = false;             // pageable source: through the pinned ring
  int device = 0;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ev;     // one per block, recorded on copy_stream after the block's DMA
  std::vector<PinnedSlot> slots;
  int nworkers = 0;
  std::vector<std::thread> workers;
  std::vector<std::atomic<int>> arrived;   // workers done with their stripe of block b
  std::mutex mu;
  std::condition_variable cv;
  int issued = 0;                  // blocks whose DMA and event record have been enqueued (in order)
  cudaError_t error = cudaSuccess;
  // side payload (the targets): one contiguous buffer that travels on the SAME copy stream right behind block 0 --
  // on a stream of its own it was served after the whole row transfer (the copy engine drained the rows' queue first)
  const char* side_src = nullptr;
  char* side_dst = nullptr;
  size_t side_bytes = 0;
  bool side_pinned = false;
  PinnedSlot side_slot;
  cudaEvent_t side_ev = nullptr;

  int64_t rows_of(int b) const { return std::min(block_rows, rows - (int64_t)b * block_rows); }

  cudaError_t enqueue_block(int b, const void* from) {
    const int64_t r = rows_of(b);
    char* to = dst + (size_t)b * block_rows * dst_pitch;
    cudaError_t e;
    if (dst_pitch == row_bytes) e = cudaMemcpyAsync(to, from, (size_t)r * row_bytes, cudaMemcpyHostToDevice, copy_stream);
    else e = cudaMemcpy2DAsync(to, dst_pitch, from, row_bytes, row_bytes, (size_t)r, cudaMemcpyHostToDevice, copy_stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(ev[b], copy_stream)) != cudaSuccess) return e;
    if (b == 0 && side_bytes) {
      const void* sfrom = side_src;
      if (!side_pinned) { host_copy_streaming(side_slot.ptr, side_src, side_bytes); sfrom = side_slot.ptr; }
      if ((e = cudaMemcpyAsync(side_dst, sfrom, side_bytes, cudaMemcpyHostToDevice, copy_stream)) != cudaSuccess) return e;
      e = cudaEventRecord(side_ev, copy_stream);
    }
    return e;
  }

  void mark_issued(int b, cudaError_t e) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return issued == b; });          // events are recorded in block order
    if (e != cudaSuccess && error == cudaSuccess) error = e;
    issued = b + 1;
    lk.unlock();
    cv.notify_all();
  }

  void worker(int w) {
    cudaSetDevice(device);
    const int nslots = (int)slots.size();
    for (int b = 0; b < nblocks; ++b) {
      if (b >= nslots) {
        // the slot still feeds the DMA of block b - nslots: wait until that has been enqueued, then until it is done
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return issued > b - nslots; });
        }
        cudaEventSynchronize(ev[b - nslots]);
      }
      const size_t bytes = (size_t)rows_of(b) * row_bytes;
      const size_t stripe = ((bytes + nworkers - 1) / nworkers + 63) & ~(size_t)63;
      const size_t lo = std::min(bytes, stripe * w), hi = std::min(bytes, lo + stripe);
      char* slot = (char*)slots[b % nslots].ptr;
      if (hi > lo) host_copy_streaming(slot + lo, src + (size_t)b * block_rows * row_bytes + lo, hi - lo);
      if (arrived[b].fetch_add(1, std::memory_order_acq_rel) == nworkers - 1) {
        // last stripe in: this worker hands the block to the copy engine
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return issued == b; });
        }
        mark_issued(b, enqueue_block(b, slot));
      }
    }
  }
};

static bool is_pinned(const void* p) {
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
                      (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
  cudaGetLastError();                                   // an unregistered pointer may leave an error code behind
  return pinned;
}

H2DTransfer* h2d_open(const void* src, void* dst, int64_t rows, size_t row_bytes, size_t dst_pitch, int64_t block_rows,
                      int threads, int slots, cudaStream_t order_after, const void* side_src, void* side_dst,
                      size_t side_bytes, cudaError_t* err) {
  *err = cudaSuccess;
  auto* t = new H2DTransfer();
  t->src = (const char*)src; t->dst = (char*)dst; t->rows = rows; t->row_bytes = row_bytes; t->dst_pitch = dst_pitch;
  t->block_rows = std::max<int64_t>(1, block_rows);
  t->nblocks = (int)((rows + t->block_rows - 1) / t->block_rows);
  cudaError_t e = cudaGetDevice(&t->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { *err = e; delete t; return nullptr; }
  // the destination may have been handed out by a stream-ordered allocator: no copy before the work already
  // enqueued on the consumer's stream
  cudaEvent_t fence;
  if ((e = cudaEventCreateWithFlags(&fence, cudaEventDisableTiming)) == cudaSuccess) {
    if ((e = cudaEventRecord(fence, order_after)) == cudaSuccess) e = cudaStreamWaitEvent(t->copy_stream, fence, 0);
    cudaEventDestroy(fence);
  }
  t->ev.resize(t->nblocks);
  for (int b = 0; b < t->nblocks && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&t->ev[b], cudaEventDisableTiming);
  if (e != cudaSuccess) { *err = e; h2d_close(t); return nullptr; }
  const bool pinned = is_pinned(src);
  if (side_src && side_dst && side_bytes) {
    t->side_src = (const char*)side_src; t->side_dst = (char*)side_dst; t->side_bytes = side_bytes;
    t->side_pinned = is_pinned(side_src);
    if ((e = cudaEventCreateWithFlags(&t->side_ev, cudaEventDisableTiming)) != cudaSuccess) { *err = e; h2d_close(t); return nullptr; }
    if (!t->side_pinned) {
      t->side_slot = g_pool.take(side_bytes);
      if (!t->side_slot.ptr) { *err = cudaErrorMemoryAllocation; h2d_close(t); return nullptr; }
    }
  }
  // (a small pageable source is not worth threads: the driver's own bounce buffer moves it in microseconds)
  t->staged = !pinned && (size_t)rows * row_bytes > ((size_t)4 << 20);
  // A pinned source is enqueued by h2d_wait itself, a bounded number of rows ahead of the consumer: the copy engine
  // works through its queue in order, so a 2 GB transfer enqueued at once would hold up every small upload the
  // caller makes next (hyper-parameters, inducing inputs) for the whole transfer.
  if (!t->staged) return t;
  const size_t slot_bytes = (size_t)t->block_rows * row_bytes;
  const int nslots = std::max(2, std::min(slots, t->nblocks));
  for (int i = 0; i < nslots; ++i) {
    PinnedSlot s = g_pool.take(slot_bytes);
    if (!s.ptr) { *err = cudaErrorMemoryAllocation; h2d_close(t); return nullptr; }
    t->slots.push_back(s);
  }
  t->nworkers = std::max(1, threads);
  t->arrived = std::vector<std::atomic<int>>(t->nblocks);
  for (auto& a : t->arrived) a.store(0);
  for (int w = 0; w < t->nworkers; ++w) t->workers.emplace_back(&H2DTransfer::worker, t, w);
  return t;
}

cudaError_t h2d_wait(H2DTransfer* t, int64_t upto_row, int64_t ahead_rows, cudaStream_t consumer) {
  if (upto_row <= 0) return cudaSuccess;
  const int last = (int)std::min<int64_t>(t->nblocks - 1, (upto_row - 1) / t->block_rows);
  if (!t->staged) {
    // single consumer thread: enqueue up to ahead_rows beyond what is asked for
    const int64_t want = std::min(t->rows, upto_row + std::max<int64_t>(0, ahead_rows));
    const int upto_block = (int)std::min<int64_t>(t->nblocks, (want + t->block_rows - 1) / t->block_rows);
    while (t->issued < upto_block) {
      const int b = t->issued;
      cudaError_t e = t->enqueue_block(b, t->src + (size_t)b * t->block_rows * t->row_bytes);
      if (e != cudaSuccess) { t->error = e; return e; }
      t->issued = b + 1;
    }
    return cudaStreamWaitEvent(consumer, t->ev[last], 0);
  }
  {
    std::unique_lock<std::mutex> lk(t->mu);
    t->cv.wait(lk, [&] { return t->issued > last || t->error != cudaSuccess; });
    if (t->error != cudaSuccess) return t->error;
  }
  return cudaStreamWaitEvent(consumer, t->ev[last], 0);   // the copy stream is in order: earlier blocks are covered
}

cudaError_t h2d_wait_side(H2DTransfer* t, cudaStream_t consumer) {
  if (!t->side_bytes) return cudaSuccess;
  cudaError_t e = h2d_wait(t, 1, 0, consumer);            // the side payload is enqueued with block 0
  if (e != cudaSuccess) return e;
  return cudaStreamWaitEvent(consumer, t->side_ev, 0);
}

bool h2d_staged(const H2DTransfer* t) { return t->staged; }

void h2d_close(H2DTransfer* t) {
  if (!t) return;
  for (auto& th : t->workers) if (th.joinable()) th.join();
  // slots go back to the pool only once their last DMA has drained
  if (!t->slots.empty() && t->copy_stream) cudaStreamSynchronize(t->copy_stream);
  for (auto& s : t->slots) g_pool.give(s);
  if (t->side_slot.ptr) { if (t->copy_stream) cudaStreamSynchronize(t->copy_stream); g_pool.give(t->side_slot); }
  if (t->side_ev) cudaEventDestroy(t->side_ev);
  for (auto& e : t->ev) if (e) cudaEventDestroy(e);
  if (t->copy_stream) {
    cudaStreamSynchronize(t->copy_stream);              // a pinned source must stay valid until its copies are done
    cudaStreamDestroy(t->copy_stream);
  }
  delete t;
}

}  // namespace edrgp
