// Host-only test of the per-layer graph engine's executor (csrc/executor.cu) on the emulated CUDA runtime
// (fake_cuda/cuda_runtime.h, including its stream-capture emulation).
//
// The step kernels of a slot are replaced by one function captured into the slot's graph (and into its group's graph,
// like Worker.prepare() does): it checksums the batch it finds in the slot's device buffers and writes the step result
// into the slot's pinned result word. The native loop `dm_exec_run` (head singles -> groups of U steps fed by gather
// threads through a ring of pinned buffers -> tail singles) must hand every step exactly the batch the `next_batch`
// sequence prescribes (/root/reference/distributed_server-basic.py:111), also with an early stop; the single-step
// submit / result API is checked too.
#include <cuda_runtime.h>
#include <stdio.h>

#include <cmath>
#include <vector>

#include "loader.h"
#include "protocol.h"

extern "C" {
void* dm_loader_create(const void*, const void*, size_t, size_t, size_t, size_t, size_t, int, uint64_t, int);
void dm_loader_next(void*, void*, void*);
uint64_t dm_loader_epochs(void*);
void dm_loader_destroy(void*);
int dm_exec_create(int, int, int, int, size_t, size_t, void**);
int dm_exec_slot_info(void*, int, void**, void**, void**, void**, void**);
void* dm_exec_capture_stream(void*, int);
int dm_exec_begin_capture(void*, int);
int dm_exec_end_capture(void*, int, int);
int dm_exec_begin_group_capture(void*, int);
int dm_exec_end_group_capture(void*, int);
int dm_exec_acquire_slot(void*, int*);
int dm_exec_submit(void*, const void*, const void*, uint64_t*);
int dm_exec_result(void*, uint64_t, void*, int);
int dm_exec_run(void*, void*, uint64_t, void*, uint32_t, uint64_t*);
int dm_exec_drain(void*);
uint64_t dm_exec_submitted(void*);
int dm_exec_destroy(void*);
const char* dm_exec_last_error();
}

namespace {
std::atomic<uint32_t> g_global_step{0};
float checksum(const float* p, size_t n) {
  double a = 0;
  for (size_t i = 0; i < n; ++i) a += p[i] * static_cast<double>((i % 7) + 1);
  return static_cast<float>(a);
}
}  // namespace

#define CHECK(c)                                                                                        \
  do {                                                                                                  \
    if (!(c)) {                                                                                         \
      fprintf(stderr, "CHECK FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #c, dm_exec_last_error()); \
      return 1;                                                                                         \
    }                                                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const int nslots = argc > 1 ? atoi(argv[1]) : 8;
  const int lanes = argc > 2 ? atoi(argv[2]) : 4;
  const int U = argc > 3 ? atoi(argv[3]) : 2;
  const uint64_t total = argc > 4 ? atoll(argv[4]) : 600;
  const size_t n = 700;
  const int I = 24, C = 10, B = 32;
  const size_t x_bytes = static_cast<size_t>(B) * I * 4, y_bytes = static_cast<size_t>(B) * C * 4;
  std::vector<float> images(n * I), labels(n * C, 0.f);
  for (size_t r = 0; r < n; ++r) {
    for (int k = 0; k < I; ++k) images[r * I + k] = static_cast<float>(r) + 0.001f * k;
    labels[r * C + r % C] = static_cast<float>(1 + r % 5);
  }
  void* ld = dm_loader_create(images.data(), labels.data(), n, I * 4, C * 4, I * 4, C * 4, B, 99, 1);
  void* ref = dm_loader_create(images.data(), labels.data(), n, I * 4, C * 4, I * 4, C * 4, B, 99, 1);
  void* ex = nullptr;
  CHECK(dm_exec_create(0, nslots, lanes, U, x_bytes, y_bytes, &ex) == 0);
  // ---- capture: one "kernel" per slot, in the slot's own graph and in its group's graph ----
  struct SlotBuf { void *xd, *yd, *res, *xs, *ys; };
  std::vector<SlotBuf> sb(nslots);
  for (int s = 0; s < nslots; ++s) CHECK(dm_exec_slot_info(ex, s, &sb[s].xd, &sb[s].yd, &sb[s].res, &sb[s].xs, &sb[s].ys) == 0);
  auto kernel_of = [&](int s) {
    const SlotBuf b = sb[s];
    return [b, I, C, B] {
      fakecuda::random_delay();
      dm::StepResult r;
      r.loss = checksum(static_cast<const float*>(b.xd), static_cast<size_t>(B) * I);
      r.correct = static_cast<uint32_t>(std::lround(checksum(static_cast<const float*>(b.yd), static_cast<size_t>(B) * C)));
      r.global_step = g_global_step.fetch_add(1) + 1;
      r.seq = r.global_step;
      *static_cast<dm::StepResult*>(b.res) = r;
    };
  };
  for (int s = 0; s < nslots; ++s) {
    CHECK(dm_exec_begin_capture(ex, s) == 0);
    static_cast<cudaStream_t>(dm_exec_capture_stream(ex, s))->push(kernel_of(s));
    CHECK(dm_exec_end_capture(ex, s, 1) == 0);
  }
  if (U > 1) {
    for (int g = 0; g < nslots / U; ++g) {
      CHECK(dm_exec_begin_group_capture(ex, g) == 0);
      for (int u = 0; u < U; ++u) static_cast<cudaStream_t>(dm_exec_capture_stream(ex, g * U + u))->push(kernel_of(g * U + u));
      CHECK(dm_exec_end_group_capture(ex, g) == 0);
    }
  }
  std::vector<float> bx(B * I), by(B * C);
  auto expect_next = [&](const dm::StepResult& r, const char* what, uint64_t step) -> bool {
    dm_loader_next(ref, bx.data(), by.data());
    const float el = checksum(bx.data(), bx.size());
    const uint32_t ec = static_cast<uint32_t>(std::lround(checksum(by.data(), by.size())));
    if (r.loss != el || r.correct != ec) {
      fprintf(stderr, "MISMATCH (%s) at step %llu: x %.3f vs %.3f, y %u vs %u\n", what, (unsigned long long)step, r.loss, el,
              r.correct, ec);
      return false;
    }
    return true;
  };
  // ---- single-step API: acquire -> fill staging -> submit -> result ----
  uint64_t done_total = 0;
  for (int k = 0; k < 3; ++k) {
    int slot = -1;
    CHECK(dm_exec_acquire_slot(ex, &slot) == 0);
    dm_loader_next(ld, sb[slot].xs, sb[slot].ys);
    uint64_t ticket = 0;
    CHECK(dm_exec_submit(ex, sb[slot].xs, sb[slot].ys, &ticket) == 0);
    dm::StepResult r;
    CHECK(dm_exec_result(ex, ticket, &r, 1) == 0);
    CHECK(expect_next(r, "submit", done_total));
    ++done_total;
  }
  // ---- native loop, runs of assorted lengths (head singles / groups / tail singles all occur) ----
  const uint64_t run_sizes[] = {5, 20, 1, 50, 16, 37, 100, 4, 211};
  size_t k = 0;
  while (done_total < total) {
    const uint64_t want = run_sizes[k++ % (sizeof(run_sizes) / sizeof(run_sizes[0]))];
    std::vector<dm::StepResult> out(want);
    uint64_t n_done = 0;
    CHECK(dm_exec_run(ex, ld, want, out.data(), 0, &n_done) == 0);
    CHECK(n_done == want);
    for (uint64_t s = 0; s < want; ++s) CHECK(expect_next(out[s], "run", done_total + s));
    done_total += want;
  }
  CHECK(dm_loader_epochs(ld) == dm_loader_epochs(ref));
  CHECK(dm_exec_submitted(ex) == done_total);
  // ---- early stop ----
  {
    const uint32_t stop_at = g_global_step.load() + 10;
    std::vector<dm::StepResult> out(200);
    uint64_t n_done = 0;
    CHECK(dm_exec_run(ex, ld, 200, out.data(), stop_at, &n_done) == 0);
    CHECK(n_done >= 10 && n_done < 200);
    uint32_t mx = 0;
    for (uint64_t s = 0; s < n_done; ++s) mx = std::max(mx, out[s].global_step);
    CHECK(mx >= stop_at);
    uint64_t n2 = 0;
    CHECK(dm_exec_run(ex, ld, 9, out.data(), 0, &n2) == 0 && n2 == 9);
  }
  printf("OK nslots=%d lanes=%d U=%d steps=%llu epochs=%llu\n", nslots, lanes, U, (unsigned long long)done_total,
         (unsigned long long)dm_loader_epochs(ld));
  CHECK(dm_exec_destroy(ex) == 0);
  dm_loader_destroy(ld);
  dm_loader_destroy(ref);
  return 0;
}
