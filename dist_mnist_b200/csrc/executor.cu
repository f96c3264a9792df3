// Native worker-side step executor and batch loader.
//
// The reference's hot loop is `next_batch(32)` + `sess.run([train_op, loss, global_step], feed_dict)`
// (/root/reference/distributed_server-basic.py:110-113): a host-driven iteration that feeds a numpy batch,
// runs the step and reads loss / global_step back. Here one training step is a chain of kernels inside a CUDA
// graph (captured once from the kernel launch plans); the executor pipelines  gather (helper threads) -> H2D
// copy (copy stream) -> graph launch (run streams) -> result (written by the step's kernel straight into pinned
// host memory)  over a ring of slots. `lanes` steps of the worker are in flight at once (asynchronous SGD); U
// consecutive slots form a group whose steps are parallel chains of one graph, so the native loops pay one
// packed input transfer and one graph launch per U steps, and the host never blocks on what it just submitted.
//
//   BatchLoader : TF `DataSet.next_batch` semantics (shuffle per epoch, sequential batches, epoch wrap)
//                 over a host-resident dataset; plan() draws the row indices (sequential), copy() moves the rows
//                 (any thread).
//   Executor    : slot ring {device x/y, pinned staging, pinned result, graph exec, events} + groups
//                 {contiguous buffers, U-step graph, run stream}; submit / submit_group / result / run.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "loader.h"
#include "trace.h"
#include "protocol.h"

namespace {

using dm::BatchLoader;

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

struct ExecSlot {
  void* x_dev = nullptr;          // into the group's device buffer
  void* y_dev = nullptr;
  dm::StepResult* res_host = nullptr;
  uint8_t* x_stage = nullptr;     // into the group's pinned staging buffer
  uint8_t* y_stage = nullptr;
  cudaGraph_t graph = nullptr;    // this slot's step alone (submit / step API, loop remainders)
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t in_ready = nullptr, done = nullptr;
  cudaEvent_t wait_ev = nullptr;  // event that marks the in-flight step complete (own `done` or the group's)
  uint64_t ticket = 0;
  bool in_flight = false;
};

// `U` consecutive slots form a group: their device and staging buffers are contiguous and one CUDA graph runs
// all U steps as parallel chains, so a native loop pays one input transfer + one graph launch per U steps.
struct ExecGroup {
  uint8_t* dev_base = nullptr;
  uint8_t* stage_base = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t in_ready = nullptr, done = nullptr;
  cudaStream_t stream = nullptr;  // every launch that touches this group's slots is ordered on this stream
};

constexpr size_t kHistory = 1 << 16;

struct Executor {
  int device = 0;
  int U = 1;                          // steps per group graph
  std::vector<cudaStream_t> gstream;  // lanes / U run streams; group g runs on gstream[g % size]
  std::vector<cudaStream_t> cap;      // U capture streams (the parallel chains of a group graph)
  std::vector<cudaEvent_t> cap_ev;    // fork / join events of the capture and of dm_exec_join / fork
  cudaStream_t copy = nullptr;
  std::vector<ExecSlot> slots;
  std::vector<ExecGroup> groups;
  size_t x_bytes = 0, y_bytes = 0, x_al = 0, slot_bytes = 0;
  uint64_t submitted = 0;  // tickets are 1-based
  std::vector<dm::StepResult> history;
  uint64_t launches = 0;
  int kernels_per_graph = 0;
  std::vector<uint8_t*> ring;  // pinned group-sized batches filled ahead by dm_exec_run's gather thread

  ExecGroup& group_of(size_t slot) { return groups[slot / U]; }

  int retire(ExecSlot& s) {
    if (!s.in_flight) return 0;
    EX_CUDA(cudaEventSynchronize(s.wait_ev));
    history[s.ticket % kHistory] = *s.res_host;
    s.in_flight = false;
    return 0;
  }
};

}  // namespace

extern "C" {

const char* dm_exec_last_error() { return g_exec_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// executor
// ---------------------------------------------------------------------------------------------
// nslots: ring depth (steps); lanes: steps in flight on the GPU at once; graph_steps (U): steps per graph launch
// in the native loops. lanes % U == 0, nslots % U == 0 and (nslots / U) % (lanes / U) == 0.
int dm_exec_create(int device, int nslots, int lanes, int graph_steps, size_t x_bytes, size_t y_bytes, void** out) {
  EX_CUDA(cudaSetDevice(device));
  const int U = graph_steps;
  if (lanes < 1 || U < 1 || lanes % U != 0 || nslots < lanes || nslots % U != 0 || (nslots / U) % (lanes / U) != 0) {
    g_exec_err = "executor: need lanes % graph_steps == 0, nslots % graph_steps == 0, (nslots/U) % (lanes/U) == 0";
    return -1;
  }
  Executor* ex = new Executor();
  ex->device = device;
  ex->U = U;
  ex->x_bytes = x_bytes;
  ex->y_bytes = y_bytes;
  ex->x_al = (x_bytes + 255) & ~size_t(255);
  ex->slot_bytes = ex->x_al + ((y_bytes + 255) & ~size_t(255));
  ex->history.resize(kHistory);
  ex->gstream.resize(lanes / U);
  for (auto& st : ex->gstream) EX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  ex->cap.resize(U);
  for (auto& st : ex->cap) EX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  ex->cap_ev.resize(std::max<size_t>(U, ex->gstream.size()) + 1);
  for (auto& e : ex->cap_ev) EX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  EX_CUDA(cudaStreamCreateWithFlags(&ex->copy, cudaStreamNonBlocking));
  ex->slots.resize(nslots);
  ex->groups.resize(nslots / U);
  for (size_t g = 0; g < ex->groups.size(); ++g) {
    ExecGroup& gr = ex->groups[g];
    const size_t bytes = ex->slot_bytes * U;
    EX_CUDA(cudaMalloc(reinterpret_cast<void**>(&gr.dev_base), bytes));
    EX_CUDA(cudaMemset(gr.dev_base, 0, bytes));
    EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&gr.stage_base), bytes, cudaHostAllocDefault));
    memset(gr.stage_base, 0, bytes);
    EX_CUDA(cudaEventCreateWithFlags(&gr.in_ready, cudaEventDisableTiming));
    EX_CUDA(cudaEventCreateWithFlags(&gr.done, cudaEventDisableTiming));
    gr.stream = ex->gstream[g % ex->gstream.size()];
    for (int u = 0; u < U; ++u) {
      ExecSlot& s = ex->slots[g * U + u];
      s.x_dev = gr.dev_base + u * ex->slot_bytes;
      s.y_dev = gr.dev_base + u * ex->slot_bytes + ex->x_al;
      s.x_stage = gr.stage_base + u * ex->slot_bytes;
      s.y_stage = s.x_stage + ex->x_al;
      // the step result is written by the head kernel straight into pinned (UVA-mapped) host memory
      EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s.res_host), sizeof(dm::StepResult),
                            cudaHostAllocMapped | cudaHostAllocPortable));
      memset(s.res_host, 0, sizeof(dm::StepResult));
      EX_CUDA(cudaEventCreateWithFlags(&s.in_ready, cudaEventDisableTiming));
      EX_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
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
void* dm_exec_compute_stream(void* h) { return static_cast<Executor*>(h)->gstream[0]; }
void* dm_exec_copy_stream(void* h) { return static_cast<Executor*>(h)->copy; }
int dm_exec_nslots(void* h) { return static_cast<int>(static_cast<Executor*>(h)->slots.size()); }
int dm_exec_graph_steps(void* h) { return static_cast<Executor*>(h)->U; }
// stream on which the kernels of `slot` must be launched between begin/end capture (single or group)
void* dm_exec_capture_stream(void* h, int slot) {
  Executor* ex = static_cast<Executor*>(h);
  return ex->cap[slot % ex->U];
}
// Make run stream 0 wait for everything submitted so far on the other run streams (stream-ordered join; used to
// close a device-timed region or to order a follow-up kernel after all in-flight steps).
int dm_exec_join(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  for (size_t l = 1; l < ex->gstream.size(); ++l) {
    EX_CUDA(cudaEventRecord(ex->cap_ev[l], ex->gstream[l]));
    EX_CUDA(cudaStreamWaitEvent(ex->gstream[0], ex->cap_ev[l], 0));
  }
  return 0;
}
// The reverse: every other run stream waits for what has been enqueued on stream 0 so far (opens a timed region).
int dm_exec_fork(void* h) {
  Executor* ex = static_cast<Executor*>(h);
  if (ex->gstream.size() > 1) {
    EX_CUDA(cudaEventRecord(ex->cap_ev[0], ex->gstream[0]));
    for (size_t l = 1; l < ex->gstream.size(); ++l) EX_CUDA(cudaStreamWaitEvent(ex->gstream[l], ex->cap_ev[0], 0));
  }
  return 0;
}

// Capture protocol, single step: begin(slot) -> Python launches the slot's kernel plans on
// dm_exec_capture_stream(slot) -> end(slot).
int dm_exec_begin_capture(void* h, int slot) {
  Executor* ex = static_cast<Executor*>(h);
  EX_CUDA(cudaSetDevice(ex->device));
  EX_CUDA(cudaStreamBeginCapture(ex->cap[slot % ex->U], cudaStreamCaptureModeRelaxed));
  return 0;
}
int dm_exec_end_capture(void* h, int slot, int kernels_in_graph) {
  Executor* ex = static_cast<Executor*>(h);
  ExecSlot& s = ex->slots.at(slot);
  // (no D2H node: the head kernel stores the 16-byte StepResult directly into s.res_host, see slot_info)
  EX_CUDA(cudaStreamEndCapture(ex->cap[slot % ex->U], &s.graph));
  EX_CUDA(cudaGraphInstantiate(&s.exec, s.graph, 0));
  ex->kernels_per_graph = kernels_in_graph;
  return 0;
}
// Group graph: begin(group) forks the U capture streams off stream 0; Python launches the plans of slot
// group * U + u on dm_exec_capture_stream(slot) for every u; end(group) joins them and instantiates one graph
// whose U step chains run concurrently.
int dm_exec_begin_group_capture(void* h, int group) {
  Executor* ex = static_cast<Executor*>(h);
  (void)group;
  EX_CUDA(cudaSetDevice(ex->device));
  EX_CUDA(cudaStreamBeginCapture(ex->cap[0], cudaStreamCaptureModeRelaxed));
  EX_CUDA(cudaEventRecord(ex->cap_ev[0], ex->cap[0]));
  for (int u = 1; u < ex->U; ++u) EX_CUDA(cudaStreamWaitEvent(ex->cap[u], ex->cap_ev[0], 0));
  return 0;
}
int dm_exec_end_group_capture(void* h, int group) {
  Executor* ex = static_cast<Executor*>(h);
  ExecGroup& gr = ex->groups.at(group);
  for (int u = 1; u < ex->U; ++u) {
    EX_CUDA(cudaEventRecord(ex->cap_ev[u], ex->cap[u]));
    EX_CUDA(cudaStreamWaitEvent(ex->cap[0], ex->cap_ev[u], 0));
  }
  EX_CUDA(cudaStreamEndCapture(ex->cap[0], &gr.graph));
  EX_CUDA(cudaGraphInstantiate(&gr.exec, gr.graph, 0));
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
  cudaStream_t run = ex->group_of(slot_idx).stream;
  if (ex->retire(s) != 0) return -1;
  if (x_src != nullptr) {
    // Inputs travel on the copy stream so that step i+1's transfer overlaps step i's kernels. (Issuing the
    // transfer on the step's own stream saves an event record / wait pair on the host but puts the copy's ~4 us
    // latency on the critical path: measured 14.2 -> 19.0 us/step with two steps in flight.)
    if (static_cast<const uint8_t*>(y_src) == static_cast<const uint8_t*>(x_src) + ex->x_al) {
      EX_CUDA(cudaMemcpyAsync(s.x_dev, x_src, ex->x_al + ex->y_bytes, cudaMemcpyDefault, ex->copy));  // packed x|y
    } else {
      EX_CUDA(cudaMemcpyAsync(s.x_dev, x_src, ex->x_bytes, cudaMemcpyDefault, ex->copy));
      EX_CUDA(cudaMemcpyAsync(s.y_dev, y_src, ex->y_bytes, cudaMemcpyDefault, ex->copy));
    }
    EX_CUDA(cudaEventRecord(s.in_ready, ex->copy));
    EX_CUDA(cudaStreamWaitEvent(run, s.in_ready, 0));
  }
  EX_CUDA(cudaGraphLaunch(s.exec, run));
  EX_CUDA(cudaEventRecord(s.done, run));
  s.wait_ev = s.done;
  s.ticket = t;
  s.in_flight = true;
  ex->submitted = t;
  ex->launches += ex->kernels_per_graph;
  if (ticket) *ticket = t;
  return 0;
}

// Submit the next U steps with one input transfer and one graph launch. Requires submitted % U == 0.
//   packed_src != null : U x slot_bytes laid out like a group buffer (x | pad | y | pad per step), host or device
//   else x_src/y_src   : step u reads x at x_src + u * x_pitch (x_bytes) and y at y_src + u * y_pitch (y_bytes)
// Returns the ticket of the first of the U steps.
int dm_exec_submit_group(void* h, const void* packed_src, const void* x_src, size_t x_pitch, const void* y_src,
                         size_t y_pitch, uint64_t* first_ticket) {
  Executor* ex = static_cast<Executor*>(h);
  const size_t U = ex->U;
  if (ex->submitted % U != 0) { g_exec_err = "submit_group: not aligned to a group boundary"; return -1; }
  const uint64_t t0 = ex->submitted + 1;
  const size_t slot0 = (t0 - 1) % ex->slots.size();
  ExecGroup& gr = ex->group_of(slot0);
  if (gr.exec == nullptr) { g_exec_err = "submit_group: group graph not captured"; return -1; }
  for (size_t u = 0; u < U; ++u)
    if (ex->retire(ex->slots[slot0 + u]) != 0) return -1;
  if (packed_src != nullptr) {
    EX_CUDA(cudaMemcpyAsync(gr.dev_base, packed_src, U * ex->slot_bytes, cudaMemcpyDefault, ex->copy));
  } else {
    EX_CUDA(cudaMemcpy2DAsync(gr.dev_base, ex->slot_bytes, x_src, x_pitch, ex->x_bytes, U, cudaMemcpyDefault, ex->copy));
    EX_CUDA(cudaMemcpy2DAsync(gr.dev_base + ex->x_al, ex->slot_bytes, y_src, y_pitch, ex->y_bytes, U,
                              cudaMemcpyDefault, ex->copy));
  }
  EX_CUDA(cudaEventRecord(gr.in_ready, ex->copy));
  EX_CUDA(cudaStreamWaitEvent(gr.stream, gr.in_ready, 0));
  EX_CUDA(cudaGraphLaunch(gr.exec, gr.stream));
  EX_CUDA(cudaEventRecord(gr.done, gr.stream));
  for (size_t u = 0; u < U; ++u) {
    ExecSlot& s = ex->slots[slot0 + u];
    s.wait_ev = gr.done;
    s.ticket = t0 + u;
    s.in_flight = true;
  }
  ex->submitted += U;
  ex->launches += static_cast<uint64_t>(ex->kernels_per_graph) * U;
  if (first_ticket) *first_ticket = t0;
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
      cudaError_t q = cudaEventQuery(s.wait_ev);
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
  for (cudaStream_t c : ex->gstream) EX_CUDA(cudaStreamSynchronize(c));
  return 0;
}

namespace {
// copies finished results [harvested, upto) of a run that started at ticket `first` into out[]
void harvest(Executor* ex, dm::StepResult* out, uint64_t first, uint64_t& harvested, uint64_t upto,
             uint32_t stop_at_global_step, bool& stop) {
  while (harvested < upto) {
    const uint64_t tk = first + harvested;
    ExecSlot& hs = ex->slots[(tk - 1) % ex->slots.size()];
    if (hs.in_flight && hs.ticket == tk) break;
    if (out) {
      out[harvested] = ex->history[tk % kHistory];
      if (stop_at_global_step && out[harvested].global_step >= stop_at_global_step) stop = true;
    }
    ++harvested;
  }
}
}  // namespace

// The native train loop: n_steps x { next_batch -> pinned staging -> H2D -> step graph -> result }. Steps are
// issued U at a time (one packed H2D transfer + one graph launch per group, see dm_exec_submit_group); a gather
// thread runs the loader up to kAhead groups ahead into a ring of pinned buffers while this thread submits, so
// the per-step host cost is max(gather, submit / U). Results of all steps are written to out_results[n_steps]
// (drained at the end). stop_at_global_step > 0 ends the loop early once a completed step reports
// global_step >= that value (StopAtStepHook semantics, reference DS:101); the number of steps actually
// submitted is returned in *n_done (batches gathered ahead of an early stop are dropped).
int dm_exec_run(void* h, void* loader, uint64_t n_steps, void* out_results, uint32_t stop_at_global_step,
                uint64_t* n_done) {
  Executor* ex = static_cast<Executor*>(h);
  BatchLoader* ld = static_cast<BatchLoader*>(loader);
  dm::StepResult* out = static_cast<dm::StepResult*>(out_results);
  dm::NvtxRange nvtx_run("dm.exec.run");
  const size_t U = ex->U;
  const size_t G = ex->groups.size();
  const uint64_t first = ex->submitted + 1;
  uint64_t harvested = 0;  // results [0, harvested) are final
  uint64_t i = 0;
  bool stop = false;
  auto single = [&]() -> int {
    ExecSlot& s = ex->slots[ex->submitted % ex->slots.size()];
    if (ex->retire(s) != 0) return -1;  // staging buffer of this slot is free again
    ld->next(s.x_stage, s.y_stage);
    if (dm_exec_submit(h, s.x_stage, s.y_stage, nullptr) != 0) return -1;
    ++i;
    harvest(ex, out, first, harvested, i, stop_at_global_step, stop);
    return 0;
  };
  // head: single steps up to the next group boundary
  while (i < n_steps && !stop && ex->submitted % U != 0)
    if (single() != 0) return -1;
  const uint64_t n_groups = (stop || ex->groups[0].exec == nullptr) ? 0 : (n_steps - i) / U;
  if (n_groups > 0) {
    constexpr size_t kAhead = 4;
    const size_t R = G + kAhead;
    const size_t gbytes = ex->slot_bytes * U;
    if (ex->ring.size() != R) {
      for (uint8_t* b : ex->ring) cudaFreeHost(b);
      ex->ring.assign(R, nullptr);
      for (auto& b : ex->ring) {
        EX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b), gbytes, cudaHostAllocDefault));
        memset(b, 0, gbytes);
      }
    }
    // Ring buffer j % R was last read by the H2D copy of group j - R, complete once that group is retired;
    // submitting group t retires group t - G, so a gather thread may fill group j once j < submitted + R - G.
    // Gathering 32 random 3 KB rows is DRAM-latency bound (~20 us per batch on one core), so several threads
    // share it: a thread claims the next group under a mutex, draws its row indices there (the loader's cursor
    // and shuffle stay sequential, i.e. the batch sequence is independent of the thread count) and copies the
    // rows outside the lock.
    std::atomic<uint64_t> subm{0};
    std::vector<std::atomic<uint64_t>> ready(R);   // ready[j % R] == j + 1: group j is in its ring buffer
    for (auto& r : ready) r.store(0, std::memory_order_relaxed);
    std::atomic<bool> quit{false};
    std::mutex plan_mu;
    uint64_t next_group = 0;
    int n_threads = 4;
    if (const char* e = getenv("DM_GATHER_THREADS")) n_threads = std::max(1, atoi(e));
    n_threads = static_cast<int>(std::min<uint64_t>(n_threads, std::min<uint64_t>(kAhead, n_groups)));
    auto gather_fn = [&] {
      std::vector<uint32_t> idx(static_cast<size_t>(U) * ld->batch);
      for (;;) {
        uint64_t j;
        {
          std::unique_lock<std::mutex> lk(plan_mu);
          for (;;) {
            if (quit.load(std::memory_order_relaxed) || next_group >= n_groups) return;
            if (next_group < subm.load(std::memory_order_acquire) + kAhead) break;
            lk.unlock();
            std::this_thread::yield();
            lk.lock();
          }
          j = next_group++;
          for (size_t u = 0; u < U; ++u) ld->plan(idx.data() + u * ld->batch);
        }
        uint8_t* b = ex->ring[j % R];
        for (size_t u = 0; u < U; ++u)
          ld->copy(idx.data() + u * ld->batch, b + u * ex->slot_bytes, b + u * ex->slot_bytes + ex->x_al);
        ready[j % R].store(j + 1, std::memory_order_release);
      }
    };
    std::vector<std::thread> gatherers;
    for (int t = 0; t < n_threads; ++t) gatherers.emplace_back(gather_fn);
    int rc = 0;
    for (uint64_t g = 0; g < n_groups && !stop; ++g) {
      while (ready[g % R].load(std::memory_order_acquire) != g + 1) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      if (dm_exec_submit_group(h, ex->ring[g % R], nullptr, 0, nullptr, 0, nullptr) != 0) { rc = -1; break; }
      subm.store(g + 1, std::memory_order_release);
      i += U;
      harvest(ex, out, first, harvested, i, stop_at_global_step, stop);
    }
    quit.store(true, std::memory_order_relaxed);
    for (auto& t : gatherers) t.join();
    if (rc != 0) return rc;
  }
  // tail: the remaining (< U) steps one by one
  while (i < n_steps && !stop)
    if (single() != 0) return -1;
  if (dm_exec_drain(h) != 0) return -1;
  for (; harvested < i; ++harvested) out[harvested] = ex->history[(first + harvested) % kHistory];
  if (n_done) *n_done = i;
  return 0;
}

// Native loop over a *device-resident* dataset (benchmark "kernel-side" number): step i takes the batch of
// `batch_rows` consecutive rows starting at ((start + i) * batch_rows) % n_rows; inputs reach the slot buffers
// with device-to-device copies on the copy stream (overlapped with the previous steps' kernels) — one strided
// copy per group of U steps whenever the group's rows are contiguous in the dataset.
int dm_exec_run_resident(void* h, uint64_t n_steps, const void* x_base, const void* y_base, size_t x_row_bytes,
                         size_t y_row_bytes, uint64_t n_rows, uint64_t batch_rows, uint64_t start) {
  Executor* ex = static_cast<Executor*>(h);
  dm::NvtxRange nvtx_run("dm.exec.run_resident");
  const uint8_t* xb = static_cast<const uint8_t*>(x_base);
  const uint8_t* yb = static_cast<const uint8_t*>(y_base);
  const size_t U = ex->U;
  const bool can_group = U > 1 && ex->groups[0].exec != nullptr && batch_rows * x_row_bytes == ex->x_bytes &&
                         batch_rows * y_row_bytes == ex->y_bytes;
  uint64_t i = 0;
  while (i < n_steps) {
    const uint64_t r = ((start + i) * batch_rows) % n_rows;
    if (can_group && ex->submitted % U == 0 && n_steps - i >= U && r + U * batch_rows <= n_rows) {
      if (dm_exec_submit_group(h, nullptr, xb + r * x_row_bytes, batch_rows * x_row_bytes, yb + r * y_row_bytes,
                               batch_rows * y_row_bytes, nullptr) != 0)
        return -1;
      i += U;
    } else {
      if (dm_exec_submit(h, xb + r * x_row_bytes, yb + r * y_row_bytes, nullptr) != 0) return -1;
      ++i;
    }
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
    cudaFreeHost(s.res_host);
    cudaEventDestroy(s.in_ready);
    cudaEventDestroy(s.done);
  }
  for (auto& g : ex->groups) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    cudaFree(g.dev_base);
    cudaFreeHost(g.stage_base);
    cudaEventDestroy(g.in_ready);
    cudaEventDestroy(g.done);
  }
  for (uint8_t* b : ex->ring) cudaFreeHost(b);
  for (cudaStream_t c : ex->gstream) cudaStreamDestroy(c);
  for (cudaStream_t c : ex->cap) cudaStreamDestroy(c);
  for (cudaEvent_t e : ex->cap_ev) cudaEventDestroy(e);
  cudaStreamDestroy(ex->copy);
  delete ex;
  return 0;
}

}  // extern "C"
