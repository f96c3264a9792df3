// Emulated CUDA runtime for host-only tests of the native executors (tests/test_fexec_emulated.py).
//
// Only what csrc/fused_exec.cu uses. Streams are real threads that execute their operations in order and
// asynchronously to the caller, events are tickets, `cudaMemcpyAsync` reads its source when the operation *executes*
// (as a DMA engine does) after a random delay — so an executor that reuses a staging buffer, an epoch buffer or a ring
// slot before the stream is done with it produces wrong data in the test, exactly as it would on the GPU.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <random>
#include <set>
#include <thread>
#include <vector>

typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3,
                      cudaMemcpyDefault = 4 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum : unsigned { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2 };
enum : unsigned { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

namespace fakecuda {

inline int max_delay_us() {
  static int v = [] { const char* e = getenv("FAKE_CUDA_DELAY_US"); return e ? atoi(e) : 50; }();
  return v;
}
inline void random_delay() {
  const int m = max_delay_us();
  if (m <= 0) return;
  thread_local std::mt19937 rng(std::random_device{}());
  const int us = static_cast<int>(rng() % static_cast<unsigned>(m + 1));
  const auto t0 = std::chrono::steady_clock::now();
  while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(us)) {}
}

struct Event {
  std::mutex mu;
  std::condition_variable cv;
  uint64_t recorded = 0, completed = 0;
  std::chrono::steady_clock::time_point when;
  uint64_t captured_in = 0;   // id of the stream capture this event was last recorded in (0: recorded for real)
};

// Stream capture: operations issued to a capturing stream (the origin and every stream that joined through an event
// wait) are appended to the graph in call order instead of being executed. Call order is one valid topological order
// of the captured dependency graph, so a launch replays the operations sequentially on the launching stream.
struct Graph {
  std::vector<std::function<void()>> ops;
};
struct Stream;
struct Capture {
  uint64_t id = 0;
  Graph* graph = nullptr;
  std::set<Stream*> members;
};
inline Capture*& current_capture() {
  static Capture* c = nullptr;
  return c;
}
inline uint64_t& capture_counter() {
  static uint64_t n = 0;
  return n;
}

struct Stream {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  uint64_t enq = 0, done = 0;
  bool quit = false;
  std::thread th;
  Stream() : th([this] { loop(); }) {}
  ~Stream() {
    { std::lock_guard<std::mutex> lk(mu); quit = true; }
    cv.notify_all();
    th.join();
  }
  bool capturing() { Capture* c = current_capture(); return c != nullptr && c->members.count(this) != 0; }
  void push(std::function<void()> f) {
    if (capturing()) { current_capture()->graph->ops.push_back(std::move(f)); return; }
    { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(f)); ++enq; }
    cv.notify_all();
  }
  void loop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return quit || !q.empty(); });
        if (q.empty()) return;
        f = std::move(q.front());
        q.pop_front();
      }
      f();
      { std::lock_guard<std::mutex> lk(mu); ++done; }
      cv.notify_all();
    }
  }
  void sync() {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t target = enq;
    cv.wait(lk, [&] { return done >= target; });
  }
  bool idle() {
    std::lock_guard<std::mutex> lk(mu);
    return done >= enq;
  }
};

}  // namespace fakecuda

typedef fakecuda::Stream* cudaStream_t;
typedef fakecuda::Event* cudaEvent_t;
typedef fakecuda::Graph* cudaGraph_t;
typedef fakecuda::Graph* cudaGraphExec_t;

inline const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : e == cudaErrorNotReady ? "cudaErrorNotReady" : "cudaErrorFake"; }
inline const char* cudaGetErrorString(cudaError_t e) { return cudaGetErrorName(e); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new fakecuda::Stream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { s->sync(); return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t s) { return s->idle() ? cudaSuccess : cudaErrorNotReady; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new fakecuda::Event(); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  if (s->capturing()) { e->captured_in = fakecuda::current_capture()->id; return cudaSuccess; }
  e->captured_in = 0;
  uint64_t ticket;
  { std::lock_guard<std::mutex> lk(e->mu); ticket = ++e->recorded; }
  s->push([e, ticket] {
    { std::lock_guard<std::mutex> lk(e->mu); if (e->completed < ticket) e->completed = ticket; e->when = std::chrono::steady_clock::now(); }
    e->cv.notify_all();
  });
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  std::unique_lock<std::mutex> lk(e->mu);
  const uint64_t t = e->recorded;
  e->cv.wait(lk, [&] { return e->completed >= t; });
  return cudaSuccess;
}
inline cudaError_t cudaEventQuery(cudaEvent_t e) {
  std::lock_guard<std::mutex> lk(e->mu);
  return e->completed >= e->recorded ? cudaSuccess : cudaErrorNotReady;
}
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->when - a->when).count();
  return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  if (fakecuda::Capture* c = fakecuda::current_capture()) {
    if (e->captured_in == c->id) { c->members.insert(s); return cudaSuccess; }   // fork / join inside the capture
  }
  uint64_t ticket;
  { std::lock_guard<std::mutex> lk(e->mu); ticket = e->recorded; }   // the most recent record at the time of the call
  s->push([e, ticket] {
    std::unique_lock<std::mutex> lk(e->mu);
    e->cv.wait(lk, [&] { return e->completed >= ticket; });
  });
  return cudaSuccess;
}
namespace fakecuda {
// Fault injection: copies whose source the predicate selects are refused (synchronous cudaErrorInvalidValue) after
// `countdown` such copies have been let through.
inline std::function<bool(const void*)>& fail_pred() { static std::function<bool(const void*)> f; return f; }
inline int& fail_countdown() { static int n = 0; return n; }
}  // namespace fakecuda
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
  if (fakecuda::fail_pred() && fakecuda::fail_pred()(src) && fakecuda::fail_countdown()-- <= 0) return cudaErrorInvalidValue;
  s->push([dst, src, n] {
    fakecuda::random_delay();
    memcpy(dst, src, n);   // the source is read now, not when the copy was enqueued
  });
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memcpy(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t s) {
  s->push([=] {
    fakecuda::random_delay();
    for (size_t r = 0; r < height; ++r)
      memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  });
  return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) {
  if (fakecuda::current_capture() != nullptr) return cudaErrorInvalidValue;
  fakecuda::Capture* c = new fakecuda::Capture();
  c->id = ++fakecuda::capture_counter();
  c->graph = new fakecuda::Graph();
  c->members.insert(s);
  fakecuda::current_capture() = c;
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g) {
  fakecuda::Capture* c = fakecuda::current_capture();
  if (c == nullptr || c->members.count(s) == 0) return cudaErrorInvalidValue;
  *g = c->graph;
  fakecuda::current_capture() = nullptr;
  delete c;
  return cudaSuccess;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) {
  *e = new fakecuda::Graph(*g);
  return cudaSuccess;
}
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s) {
  s->push([e] {
    for (auto& op : e->ops) op();
  });
  return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
