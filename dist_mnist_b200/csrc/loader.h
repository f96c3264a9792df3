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
#include <random>
#include <vector>

namespace dm {

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

  void reshuffle() {
    if (shuffle) std::shuffle(perm.begin(), perm.end(), rng);
  }
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
  void copy(const uint32_t* idx, uint8_t* x_dst, uint8_t* y_dst) const {
    for (int r = 0; r < batch; ++r) {
      memcpy(x_dst + r * x_dst_stride, images + static_cast<size_t>(idx[r]) * x_row_bytes, x_row_bytes);
      memcpy(y_dst + r * y_dst_stride, labels + static_cast<size_t>(idx[r]) * y_row_bytes, y_row_bytes);
    }
  }
  // rows [r0, r1) of a planned batch (a batch can be gathered by several threads)
  void copy_rows(const uint32_t* idx, int r0, int r1, uint8_t* x_dst, uint8_t* y_dst) const {
    for (int r = r0; r < r1; ++r) {
      memcpy(x_dst + r * x_dst_stride, images + static_cast<size_t>(idx[r]) * x_row_bytes, x_row_bytes);
      memcpy(y_dst + r * y_dst_stride, labels + static_cast<size_t>(idx[r]) * y_row_bytes, y_row_bytes);
    }
  }
  void next(uint8_t* x_dst, uint8_t* y_dst) {
    std::vector<uint32_t> idx(batch);
    plan(idx.data());
    copy(idx.data(), x_dst, y_dst);
  }
};

}  // namespace dm
