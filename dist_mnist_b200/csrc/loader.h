// Host-side `next_batch` loader shared by the executors (executor.cu, fused_exec.cu).
//
// TF `DataSet.next_batch` semantics (the reference's `mnist.train.next_batch(32)`,
// /root/reference/distributed_server-basic.py:111): shuffle per epoch, sequential batches, an epoch boundary
// inside a batch is finished from the next epoch. plan() draws the row indices (sequential: it owns the cursor,
// the epoch counter and the shuffle), copy() moves the rows (the expensive part; safe on any thread).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <random>
#include <vector>

namespace dm {

// Row copy into a buffer the CPU will not read again (pinned staging / epoch buffers: the next reader is the DMA engine).
// Streaming (non-temporal) stores skip the read-for-ownership of every destination line and leave the caches to the
// source rows: 1.8x the rate of memcpy for 3 KB rows measured on the build box (5.5 -> 10 GB/s per thread). Falls back to
// memcpy for sizes / alignments the vector loop does not cover. Defined in loader_api.cpp (plain C++, runtime ISA check).
void copy_row_streaming(uint8_t* dst, const uint8_t* src, size_t n);
// Orders the streaming stores of this thread before whatever publishes the rows (an atomic counter, a CUDA call).
void streaming_fence();

struct BatchLoader {
  const uint8_t* images;
  const uint8_t* labels;
  size_t n;
  size_t x_row_bytes, y_row_bytes;
  size_t x_dst_stride, y_dst_stride;
  int batch;
  bool shuffle;
  std::mt19937_64 rng;
  std::vector<uint32_t> perm;
  size_t cursor = 0;
  uint64_t epochs = 0;

  // ---- epoch feed (fused_exec.cu) ----
  // TF's DataSet shuffles *physically* at an epoch boundary (`self._images = self.images[perm]`) and then hands out
  // contiguous slices; `next_batch` between boundaries copies nothing. The epoch feed is that design with the
  // boundary cost taken off the training thread: two pinned buffers hold the rows of the current / the next epoch in
  // permutation order (buffer `epoch & 1`), the next epoch's buffer is filled by the executor's helper threads while
  // the current one is being consumed, and a batch that does not straddle a boundary is a contiguous slice of pinned
  // memory that is DMA'd to the GPU as it is. The permutation of epoch e+1 is drawn ahead of time from the same
  // generator state the boundary would have used, so the sequence of batches is bit-identical with or without a feed.
  static constexpr uint64_t kNoEpoch = ~0ull;
  bool feed = false;
  uint8_t* feed_x[2] = {nullptr, nullptr};
  uint8_t* feed_y[2] = {nullptr, nullptr};
  uint64_t feed_epoch[2] = {kNoEpoch, kNoEpoch};   // epoch whose rows buffer b holds / is being filled with
  std::atomic<uint32_t> feed_rows[2];              // rows of that epoch in place
  std::shared_ptr<std::vector<uint32_t>> next_perm;   // permutation of epoch `epochs + 1`, once drawn

  BatchLoader() {
    feed_rows[0].store(0);
    feed_rows[1].store(0);
  }
  BatchLoader(const BatchLoader&) = delete;
  BatchLoader& operator=(const BatchLoader&) = delete;

  // The next epoch's permutation, drawn exactly as reshuffle() would at the boundary (the generator is used by
  // nothing else). The vector is shared with the fill job that copies the rows, so it is never modified afterwards.
  const std::shared_ptr<std::vector<uint32_t>>& draw_next_perm() {
    if (!next_perm) {
      next_perm = std::make_shared<std::vector<uint32_t>>(perm);
      if (shuffle) std::shuffle(next_perm->begin(), next_perm->end(), rng);
    }
    return next_perm;
  }
  void reshuffle() {
    if (next_perm) {
      perm = *next_perm;
      next_perm.reset();
    } else if (shuffle) {
      std::shuffle(perm.begin(), perm.end(), rng);
    }
  }
  // Rows [cursor, cursor + rows) of the current epoch are in place in its feed buffer and do not reach the boundary.
  bool feed_slice_ready(size_t rows) const {
    const int b = static_cast<int>(epochs & 1);
    return feed && cursor + rows <= n && feed_epoch[b] == epochs && feed_rows[b].load(std::memory_order_acquire) >= n;
  }
  // The current epoch's buffer is being filled (worth waiting for instead of gathering row by row).
  bool feed_fill_in_progress() const {
    const int b = static_cast<int>(epochs & 1);
    return feed && feed_epoch[b] == epochs && feed_rows[b].load(std::memory_order_acquire) < n;
  }
  // Consume `rows` rows of the current epoch without materialising their indices (rows <= n - cursor).
  void skip_rows(size_t rows) { cursor += rows; }
  // next() = plan() + copy(): plan draws the batch's row indices (sequential: it owns the cursor, the epoch
  // counter and the shuffle), copy moves the rows (the expensive part; safe to run on any thread).
  void plan(uint32_t* idx_out) {
    for (int r = 0; r < batch; ++r) {
      if (cursor == n) {  // epoch boundary inside a batch: finish it from the next epoch (TF next_batch)
        cursor = 0;
        ++epochs;
        reshuffle();
      }
      idx_out[r] = perm[cursor++];
    }
  }
  void copy(const uint32_t* idx, uint8_t* x_dst, uint8_t* y_dst) const { copy_rows(idx, 0, batch, x_dst, y_dst); }
  // rows [r0, r1) of a planned batch (a batch can be gathered by several threads). The destination is a buffer the DMA
  // engine reads next (pinned staging, epoch buffers): streaming stores.
  void copy_rows(const uint32_t* idx, int r0, int r1, uint8_t* x_dst, uint8_t* y_dst) const {
    for (int r = r0; r < r1; ++r) {
      copy_row_streaming(x_dst + r * x_dst_stride, images + static_cast<size_t>(idx[r]) * x_row_bytes, x_row_bytes);
      memcpy(y_dst + r * y_dst_stride, labels + static_cast<size_t>(idx[r]) * y_row_bytes, y_row_bytes);
    }
    streaming_fence();
  }
  // One batch into ordinary memory that the caller's CPU code reads next (Python's next_batch, the cpu backend).
  void next(uint8_t* x_dst, uint8_t* y_dst) {
    std::vector<uint32_t> idx(batch);
    plan(idx.data());
    for (int r = 0; r < batch; ++r) {
      memcpy(x_dst + r * x_dst_stride, images + static_cast<size_t>(idx[r]) * x_row_bytes, x_row_bytes);
      memcpy(y_dst + r * y_dst_stride, labels + static_cast<size_t>(idx[r]) * y_row_bytes, y_row_bytes);
    }
  }
};

}  // namespace dm
