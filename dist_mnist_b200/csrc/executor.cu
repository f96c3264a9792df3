// Native worker-side step executor and batch loader.
//
// The reference's hot loop is `next_batch(32)` + `sess.run([train_op, loss, global_step], feed_dict)`
// (/root/reference/distributed_server-basic.py:110-113): a host-driven iteration that feeds a numpy batch,
// runs the step and reads loss / global_step back. Here one training step is a CUDA graph (captured once
// from the kernel launch plans); the executor pipelines  gather -> H2D copy (copy stream) -> graph launch
// (compute stream) -> 16-byte D2H result  over a ring of slots, so step i+1's input transfer overlaps step
// i's kernels and the host never blocks on the step it just submitted.
//
//   BatchLoader : TF `DataSet.next_batch` semantics (shuffle per epoch, sequential batches, epoch wrap)
//                 over a host-resident dataset; gathers rows into a pinned staging buffer.
//   Executor    : slot ring {device x/y, pinned staging, result, graph exec, events}; submit / result / run.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "protocol.h"

namespace {

thread_local std::string g_exec_err;

int efail(const char* what, cudaError_t e) {
  g_exec_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return -1;
}
#define EX_CUDA(call)                                  \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return efail(#call, e__);  \
  } while (0)

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
  void next(uint8_t* x_dst, uint8_t* y_dst) {
    for (int r = 0; r < batch; ++r) {
      if (cursor == n) {  // epoch boundary inside a batch: finish it from the next epoch (TF next_batch)
        cursor = 0;
        ++epochs;
        reshuffle();
      }
      const size_t idx = perm[cursor++];
      memcpy(x_dst + r * x_dst_stride, images + idx * x_row_bytes, x_row_bytes);
      memcpy(y_dst + r * y_dst_stride, labels + idx * y_row_bytes, y_row_bytes);
    }
  }
};

struct ExecSlot {
  void* x_dev = nullptr;
  void* y_dev = nullptr;
  dm::StepResult* res_dev = nullptr;
  dm::StepResult* res_host = nullptr;
  uint8_t* x_stage = nullptr;
  uint8_t* y_stage = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t in_ready = nullptr, done = nullptr;
  uint64_t ticket = 0;
  bool in_flight = false;
};

constexpr size_t kHistory = 1 << 16;

struct Executor {
  int device = 0;
  // `lanes` compute streams: slot s runs on lane s % lanes, so up to `lanes` consecutive steps are in flight
  // on the GPU at once (asynchronous SGD: a step does not wait for the previous step's push).
  std::vector<cudaStream_t> compute;
  std::vector<cudaEvent_t> lane_ev;
  cudaStream_t copy = nullptr;
  std::vector<ExecSlot> slots;
  cudaStream_t lane_of(size_t slot) const { return compute[slot % compute.size()]; }
  size_t x_bytes = 0, y_bytes = 0;
  uint64_t submitted = 0;  // tickets are 1-based
  std::vector<dm::StepResult> history;
  uint64_t launches = 0;
  int kernels_per_graph = 0;
  std::vector<uint8_t*> ring;  // pinned x|y batches filled ahead by dm_exec_run's gather thread

  int retire(ExecSlot& s) {
    if (!s.in_flight) return 0;
    EX_CUDA(cudaEventSynchronize(s.done));
    history[s.ticket % kHistory] = *s.res_host;
    s.in_flight = false;
    return 0;
  }
};

}  // namespace

extern "C" {

const char* dm_exec_last_error() { return g_exec_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// batch loader
// ---------------------------------------------------------------------------------------------
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
void dm_loader_destroy(void* h) { delete static_cast<BatchLoader*>(h); }

// ---------------------------------------------------------------------------------------------
// executor
// ---------------------------------------------------------------------------------------------
int dm_exec_create(int device, int nslots, int lanes, size_t x_bytes, size_t y_bytes, void** out) {
  EX_CUDA(cudaSetDevice(device));
  if (lanes < 1 || nslots < lanes || nslots % lanes != 0) {
    g_exec_err = "executor: nslots must be a positive multiple of lanes";
    return -1;
  }
  Executor* ex = new Executor();
  ex->device = device;
  ex->x_bytes = x_bytes;
  ex->y_bytes = y_bytes;
  ex->history.resize(kHistory);
  ex->compute.resize(lanes);
  ex->lane_ev.resize(lanes);
  for (int l = 0; l < lanes; ++l) {
    EX_CUDA(cudaStreamCreateWithFlags(&ex->compute[l], cudaStreamNonBlocking));
    EX_CUDA(cudaEventCreateWithFlags(&ex->lane_ev[l], cudaEventDisableTiming));
  }
  EX_CUDA(cudaStreamCreateWithFlags(&ex->copy, cudaStreamNonBlocking));
  ex->slots.resize(nslots);
  for (auto& s : ex->slots) {
    // x and y of a slot are one allocation (device and pinned staging alike): the native loop then moves a
    // whole batch host->device with a single cudaMemcpyAsync.
    const size_t x_al = (x_bytes + 255) & ~size_t(255);
    EX_CUDA(cudaMalloc(&s.x_dev, x_al + y_bytes));
    EX_CUDA(cudaMemset(s.x_dev, 0, x_al + y_bytes));
    s.y_dev = static_cast<uint8_t*>(s.x_dev) + x_al;
    EX_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.res_dev), sizeof(dm::StepResult)));
    EX_CUDA(cudaMemset(s.res_dev, 0, sizeof(dm::StepResult)));
    EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s.res_host), sizeof(dm::StepResult),
                          cudaHostAllocMapped | cudaHostAllocPortable));  // written by the head kernel over PCIe
    EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s.x_stage), x_al + y_bytes, cudaHostAllocDefault));
    s.y_stage = s.x_stage + x_al;
    memset(s.res_host, 0, sizeof(dm::StepResult));
    memset(s.x_stage, 0, x_al + y_bytes);
    EX_CUDA(cudaEventCreateWithFlags(&s.in_ready, cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  }
  *out = ex;
  return 0;
}

int dm_exec_slot_info(void* h, int slot, void** x_dev, void** y_dev, void** res_dev, void** x_stage, void** y_stage) {
  Executor* ex = static_cast<Executor*>(h);
  ExecSlot& s = ex->slots.at(slot);
  *x_dev = s.x_dev;
  *y_dev = s.y_dev;
  // The step result is written by the head kernel straight into pinned (UVA-mapped) host memory: a 16-byte
  // posted PCIe write instead of a D2H memcpy node that would add a copy-engine round trip to every step.
  *res_dev = s.res_host;
  *x_stage = s.x_stage;
  *y_stage = s.y_stage;
  return 0;
}
void* dm_exec_compute_stream(void* h) { return static_cast<Executor*>(h)->compute[0]; }
void* dm_exec_lane_stream(void* h, int lane) { return static_cast<Executor*>(h)->compute.at(lane); }
int dm_exec_lanes(void* h) { return static_cast<int>(static_cast<Executor*>(h)->compute.size()); }
// Make lane 0 wait for everything submitted so far on the other lanes (stream-ordered join; used to close a
// device-timed region or to order a follow-up kernel after all in-flight steps).
int dm_exec_join(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  for (size_t l = 1; l < ex->compute.size(); ++l) {
    EX_CUDA(cudaEventRecord(ex->lane_ev[l], ex->compute[l]));
    EX_CUDA(cudaStreamWaitEvent(ex->compute[0], ex->lane_ev[l], 0));
  }
  return 0;
}
// The reverse: every other lane waits for what has been enqueued on lane 0 so far (opens a timed region).
int dm_exec_fork(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  if (ex->compute.size() > 1) {
    EX_CUDA(cudaEventRecord(ex->lane_ev[0], ex->compute[0]));
    for (size_t l = 1; l < ex->compute.size(); ++l) EX_CUDA(cudaStreamWaitEvent(ex->compute[l], ex->lane_ev[0], 0));
  }
  return 0;
}
void* dm_exec_copy_stream(void* h) { return static_cast<Executor*>(h)->copy; }
int dm_exec_nslots(void* h) { return static_cast<int>(static_cast<Executor*>(h)->slots.size()); }

// Capture protocol: begin -> (Python launches the step's kernel plans on the compute stream) -> end.
int dm_exec_begin_capture(void* h, int slot) {
  Executor* ex = static_cast<Executor*>(h);
  EX_CUDA(cudaSetDevice(ex->device));
  EX_CUDA(cudaStreamBeginCapture(ex->lane_of(slot), cudaStreamCaptureModeRelaxed));
  return 0;
}
int dm_exec_end_capture(void* h, int slot, int kernels_in_graph) {
  Executor* ex = static_cast<Executor*>(h);
  ExecSlot& s = ex->slots.at(slot);
  // (no D2H node: the head kernel stores the 16-byte StepResult directly into s.res_host, see slot_info)
  EX_CUDA(cudaStreamEndCapture(ex->lane_of(slot), &s.graph));
  EX_CUDA(cudaGraphInstantiate(&s.exec, s.graph, 0));
  ex->kernels_per_graph = kernels_in_graph;
  return 0;
}

// Retire (wait for) the step that last used the next slot and return that slot's index: after this call the
// slot's pinned staging buffers may be overwritten with the next batch.
int dm_exec_acquire_slot(void* h, int* slot) {
  Executor* ex = static_cast<Executor*>(h);
  const int idx = static_cast<int>(ex->submitted % ex->slots.size());
  if (ex->retire(ex->slots[idx]) != 0) return -1;
  *slot = idx;
  return 0;
}

// Submit one step. x_src / y_src: host (pinned) or device pointers of x_bytes / y_bytes, or null to reuse the
// data already resident in the slot's device buffers. Returns the 1-based ticket.
int dm_exec_submit(void* h, const void* x_src, const void* y_src, uint64_t* ticket) {
  Executor* ex = static_cast<Executor*>(h);
  const uint64_t t = ex->submitted + 1;
  const size_t slot_idx = (t - 1) % ex->slots.size();
  ExecSlot& s = ex->slots[slot_idx];
  cudaStream_t lane = ex->lane_of(slot_idx);
  if (ex->retire(s) != 0) return -1;
  if (x_src != nullptr) {
    // Inputs travel on the copy stream so that step i+1's transfer overlaps step i's kernels. (Issuing the
    // transfer on the step's own lane saves an event record / wait pair on the host but puts the copy's ~4 us
    // latency on the lane's critical path: measured 14.2 -> 19.0 us/step with two lanes.)
    cudaStream_t cs = ex->copy;
    const size_t x_al = (ex->x_bytes + 255) & ~size_t(255);
    if (static_cast<const uint8_t*>(y_src) == static_cast<const uint8_t*>(x_src) + x_al) {
      EX_CUDA(cudaMemcpyAsync(s.x_dev, x_src, x_al + ex->y_bytes, cudaMemcpyDefault, cs));  // x|y packed like a slot
    } else {
      EX_CUDA(cudaMemcpyAsync(s.x_dev, x_src, ex->x_bytes, cudaMemcpyDefault, cs));
      EX_CUDA(cudaMemcpyAsync(s.y_dev, y_src, ex->y_bytes, cudaMemcpyDefault, cs));
    }
    if (cs != lane) {
      EX_CUDA(cudaEventRecord(s.in_ready, ex->copy));
      EX_CUDA(cudaStreamWaitEvent(lane, s.in_ready, 0));
    }
  }
  EX_CUDA(cudaGraphLaunch(s.exec, lane));
  EX_CUDA(cudaEventRecord(s.done, lane));
  s.ticket = t;
  s.in_flight = true;
  ex->submitted = t;
  ex->launches += ex->kernels_per_graph;
  if (ticket) *ticket = t;
  return 0;
}

// Fetch the result of a ticket. wait != 0 blocks until it is complete. Returns 0 = ok, 1 = not ready, -1 error,
// 2 = ticket too old (fell out of the history ring).
int dm_exec_result(void* h, uint64_t ticket, void* out, int wait) {
  Executor* ex = static_cast<Executor*>(h);
  if (ticket == 0 || ticket > ex->submitted) { g_exec_err = "bad ticket"; return -1; }
  if (ex->submitted - ticket >= kHistory) return 2;
  ExecSlot& s = ex->slots[(ticket - 1) % ex->slots.size()];
  if (s.in_flight && s.ticket == ticket) {
    if (!wait) {
      cudaError_t q = cudaEventQuery(s.done);
      if (q == cudaErrorNotReady) { cudaGetLastError(); return 1; }
      if (q != cudaSuccess) return efail("cudaEventQuery", q);
    }
    if (ex->retire(s) != 0) return -1;
  }
  memcpy(out, &ex->history[ticket % kHistory], sizeof(dm::StepResult));
  return 0;
}

int dm_exec_drain(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  for (auto& s : ex->slots)
    if (ex->retire(s) != 0) return -1;
  EX_CUDA(cudaStreamSynchronize(ex->copy));
  for (cudaStream_t c : ex->compute) EX_CUDA(cudaStreamSynchronize(c));
  return 0;
}

// The native train loop: n_steps x { next_batch -> pinned staging -> H2D -> step graph -> result }. A gather
// thread runs the loader up to kAhead batches ahead into a ring of pinned x|y buffers while this thread submits,
// so the per-step host cost is max(gather, submit) instead of their sum. Results of all steps are written to
// out_results[n_steps] (drained at the end). stop_at_global_step > 0 ends the loop early once a completed step
// reports global_step >= that value (StopAtStepHook semantics, reference DS:101); the number of steps actually
// submitted is returned in *n_done (batches gathered ahead of an early stop are dropped).
int dm_exec_run(void* h, void* loader, uint64_t n_steps, void* out_results, uint32_t stop_at_global_step,
                uint64_t* n_done) {
  Executor* ex = static_cast<Executor*>(h);
  BatchLoader* ld = static_cast<BatchLoader*>(loader);
  dm::StepResult* out = static_cast<dm::StepResult*>(out_results);
  const size_t S = ex->slots.size();
  const size_t x_al = (ex->x_bytes + 255) & ~size_t(255);
  constexpr size_t kAhead = 4;
  const size_t R = S + kAhead;
  if (ex->ring.size() != R) {
    for (uint8_t* b : ex->ring) cudaFreeHost(b);
    ex->ring.assign(R, nullptr);
    for (auto& b : ex->ring) {
      EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b), x_al + ex->y_bytes, cudaHostAllocDefault));
      memset(b, 0, x_al + ex->y_bytes);
    }
  }
  // Ring buffer j % R was last read by the H2D copy of step j - R, which is complete once that step has been
  // retired; submitting step t retires step t - S, hence the gather thread may fill batch j as soon as
  // j < submitted + (R - S).
  std::atomic<uint64_t> filled{0}, submitted{0};
  std::atomic<bool> quit{false};
  std::thread gather([&] {
    for (uint64_t j = 0; j < n_steps && !quit.load(std::memory_order_relaxed); ++j) {
      while (j >= submitted.load(std::memory_order_acquire) + kAhead) {
        if (quit.load(std::memory_order_relaxed)) return;
        std::this_thread::yield();
      }
      uint8_t* b = ex->ring[j % R];
      ld->next(b, b + x_al);
      filled.store(j + 1, std::memory_order_release);
    }
  });
  const uint64_t first = ex->submitted + 1;
  uint64_t harvested = 0;  // results [0, harvested) are final
  uint64_t i = 0;
  bool stop = false;
  int rc = 0;
  for (; i < n_steps && !stop; ++i) {
    while (filled.load(std::memory_order_acquire) <= i) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    uint8_t* b = ex->ring[i % R];
    uint64_t t;
    if (dm_exec_submit(h, b, b + x_al, &t) != 0) { rc = -1; break; }
    submitted.store(i + 1, std::memory_order_release);
    // harvest whatever has been retired so far (keeps stop latency at <= nslots steps)
    while (harvested < i + 1) {
      const uint64_t tk = first + harvested;
      ExecSlot& hs = ex->slots[(tk - 1) % S];
      if (hs.in_flight && hs.ticket == tk) break;
      out[harvested] = ex->history[tk % kHistory];
      if (stop_at_global_step && out[harvested].global_step >= stop_at_global_step) stop = true;
      ++harvested;
    }
  }
  quit.store(true, std::memory_order_relaxed);
  gather.join();
  if (rc != 0) return rc;
  if (dm_exec_drain(h) != 0) return -1;
  for (; harvested < i; ++harvested) out[harvested] = ex->history[(first + harvested) % kHistory];
  if (n_done) *n_done = i;
  return 0;
}

// Native loop over a *device-resident* dataset (benchmark "kernel-side" number): step i takes the batch of
// `batch_rows` consecutive rows starting at ((start + i) * batch_rows) % n_rows; inputs reach the slot buffers
// with a device-to-device copy on the copy stream (overlapped with the previous step's kernels).
int dm_exec_run_resident(void* h, uint64_t n_steps, const void* x_base, const void* y_base, size_t x_row_bytes,
                         size_t y_row_bytes, uint64_t n_rows, uint64_t batch_rows, uint64_t start) {
  const uint8_t* xb = static_cast<const uint8_t*>(x_base);
  const uint8_t* yb = static_cast<const uint8_t*>(y_base);
  for (uint64_t i = 0; i < n_steps; ++i) {
    const uint64_t r = ((start + i) * batch_rows) % n_rows;
    if (dm_exec_submit(h, xb + r * x_row_bytes, yb + r * y_row_bytes, nullptr) != 0) return -1;
  }
  return 0;
}

uint64_t dm_exec_submitted(void* h) { return static_cast<Executor*>(h)->submitted; }
uint64_t dm_exec_kernel_launches(void* h) { return static_cast<Executor*>(h)->launches; }

int dm_exec_destroy(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  cudaSetDevice(ex->device);
  dm_exec_drain(h);
  for (auto& s : ex->slots) {
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.graph) cudaGraphDestroy(s.graph);
    cudaFree(s.x_dev);
    cudaFree(s.res_dev);
    cudaFreeHost(s.res_host);
    cudaFreeHost(s.x_stage);
    cudaEventDestroy(s.in_ready);
    cudaEventDestroy(s.done);
  }
  for (uint8_t* b : ex->ring) cudaFreeHost(b);
  for (cudaStream_t c : ex->compute) cudaStreamDestroy(c);
  for (cudaEvent_t e : ex->lane_ev) cudaEventDestroy(e);
  cudaStreamDestroy(ex->copy);
  delete ex;
  return 0;
}

}  // extern "C"
