// C API of the native `next_batch` loader (loader.h). No CUDA in this file: the loader is host code and is unit-tested
// on CPU (tests/test_properties.py, tests/test_fexec_emulated.py).
//
// Reference: `mnist.train.next_batch(32)` (/root/reference/distributed_server-basic.py:111) on TF's DataSet: shuffle
// at every epoch boundary, sequential batches, a batch that straddles the boundary is completed from the next epoch.
#include <stdlib.h>

#include <chrono>
#include <thread>

#include "loader.h"
#include "trace.h"

#if defined(__x86_64__) && !defined(__SANITIZE_THREAD__)
#include <immintrin.h>
#define DM_STREAMING_COPY 1
#endif

using dm::BatchLoader;

namespace dm {

#ifdef DM_STREAMING_COPY
namespace {
__attribute__((target("avx2"))) void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {   // dst 32 B aligned, n % 32 == 0
  for (size_t i = 0; i < n; i += 32)
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
}
void copy_nt_sse2(uint8_t* dst, const uint8_t* src, size_t n) {   // dst 16 B aligned, n % 16 == 0
  for (size_t i = 0; i < n; i += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
}
const bool g_have_avx2 = [] { __builtin_cpu_init(); return __builtin_cpu_supports("avx2") != 0; }();
const bool g_streaming = [] { const char* e = getenv("DM_STREAMING_COPY"); return e == nullptr || e[0] != '0'; }();   // A/B knob
}  // namespace

void copy_row_streaming(uint8_t* dst, const uint8_t* src, size_t n) {
  const uintptr_t d = reinterpret_cast<uintptr_t>(dst);
  if (n < 256 || !g_streaming) { memcpy(dst, src, n); return; }   // short rows: not worth bypassing the cache
  if (g_have_avx2 && (d & 31) == 0 && (n & 31) == 0) copy_nt_avx2(dst, src, n);
  else if ((d & 15) == 0 && (n & 15) == 0) copy_nt_sse2(dst, src, n);
  else memcpy(dst, src, n);
}
void streaming_fence() { _mm_sfence(); }
#else
// (ThreadSanitizer builds and non-x86 hosts: plain copies — TSan does not see streaming stores)
void copy_row_streaming(uint8_t* dst, const uint8_t* src, size_t n) { memcpy(dst, src, n); }
void streaming_fence() {}
#endif

}  // namespace dm

extern "C" {

void* dm_loader_create(const void* images, const void* labels, size_t n, size_t x_row_bytes, size_t y_row_bytes,
                       size_t x_dst_stride, size_t y_dst_stride, int batch, uint64_t seed, int shuffle) {
  BatchLoader* l = new BatchLoader();
  l->images = static_cast<const uint8_t*>(images);
  l->labels = static_cast<const uint8_t*>(labels);
  l->n = n;
  l->x_row_bytes = x_row_bytes;
  l->y_row_bytes = y_row_bytes;
  l->x_dst_stride = x_dst_stride;
  l->y_dst_stride = y_dst_stride;
  l->batch = batch;
  l->shuffle = shuffle != 0;
  l->rng.seed(seed);
  l->perm.resize(n);
  for (size_t i = 0; i < n; ++i) l->perm[i] = static_cast<uint32_t>(i);
  l->reshuffle();
  return l;
}

void dm_loader_next(void* h, void* x_dst, void* y_dst) {
  static_cast<BatchLoader*>(h)->next(static_cast<uint8_t*>(x_dst), static_cast<uint8_t*>(y_dst));
}

uint64_t dm_loader_epochs(void* h) { return static_cast<BatchLoader*>(h)->epochs; }

// Test hook: the streaming row copy on its own (any size / alignment must equal memcpy).
void dm_copy_row_streaming(void* dst, const void* src, size_t n) {
  dm::copy_row_streaming(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n);
  dm::streaming_fence();
}

// Epoch feed (loader.h): x0/y0/x1/y1 are two pairs of caller-owned pinned buffers of n rows each (x_row_bytes /
// y_row_bytes per row, densely packed). Fills the current epoch's buffer here (a few helper threads; ~170 MB for MNIST)
// and returns 1; returns 0 and leaves the loader unchanged when the layout does not allow contiguous slices
// (padded destination rows) or the dataset is too small to be worth it.
int dm_loader_enable_feed(void* h, void* x0, void* y0, void* x1, void* y1, int n_threads) {
  BatchLoader* l = static_cast<BatchLoader*>(h);
  if (l->feed) return 1;
  dm::NvtxRange nvtx("dm.loader.enable_feed");
  if (l->x_dst_stride != l->x_row_bytes || l->y_dst_stride != l->y_row_bytes) return 0;
  if (l->n < 1024 || l->n > 0x7FFFFFFFull || l->batch < 1 || static_cast<size_t>(l->batch) * 8 > l->n) return 0;
  if (!x0 || !y0 || !x1 || !y1) return 0;
  l->feed_x[0] = static_cast<uint8_t*>(x0);
  l->feed_y[0] = static_cast<uint8_t*>(y0);
  l->feed_x[1] = static_cast<uint8_t*>(x1);
  l->feed_y[1] = static_cast<uint8_t*>(y1);
  const int b = static_cast<int>(l->epochs & 1);
  const int nt = std::max(1, std::min(n_threads, 16));
  const uint32_t* idx = l->perm.data();
  std::vector<std::thread> th;
  const size_t per = (l->n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const size_t r0 = std::min(l->n, t * per), r1 = std::min(l->n, r0 + per);
    if (r0 >= r1) break;
    th.emplace_back([l, idx, b, r0, r1] {
      for (size_t r = r0; r < r1; ++r) {
        dm::copy_row_streaming(l->feed_x[b] + r * l->x_row_bytes, l->images + static_cast<size_t>(idx[r]) * l->x_row_bytes,
                               l->x_row_bytes);
        memcpy(l->feed_y[b] + r * l->y_row_bytes, l->labels + static_cast<size_t>(idx[r]) * l->y_row_bytes, l->y_row_bytes);
      }
      dm::streaming_fence();
    });
  }
  for (auto& t : th) t.join();
  l->feed_epoch[b] = l->epochs;
  l->feed_rows[b].store(static_cast<uint32_t>(l->n), std::memory_order_release);
  l->feed_epoch[b ^ 1] = BatchLoader::kNoEpoch;
  l->feed_rows[b ^ 1].store(0, std::memory_order_release);
  l->feed = true;
  return 1;
}

int dm_loader_feed_enabled(void* h) { return static_cast<BatchLoader*>(h)->feed ? 1 : 0; }

// A fill job posted to an executor's helper threads may still be writing into the feed buffers and reading the
// dataset: wait for it (the executor's threads finish every posted job, also while being shut down). If a job never
// completes (its executor is gone without having run it — cannot happen through the Python API) the loader is leaked
// rather than freed under it.
void dm_loader_destroy(void* h) {
  BatchLoader* l = static_cast<BatchLoader*>(h);
  if (l->feed) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < 2; ++b) {
      while (l->feed_epoch[b] != BatchLoader::kNoEpoch && l->feed_rows[b].load(std::memory_order_acquire) < l->n) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) return;
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    }
  }
  delete l;
}

}  // extern "C"
