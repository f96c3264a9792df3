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
#include <thread>

typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
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
};

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
  void push(std::function<void()> f) {
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
  uint64_t ticket;
  { std::lock_guard<std::mutex> lk(e->mu); ticket = e->recorded; }   // the most recent record at the time of the call
  s->push([e, ticket] {
    std::unique_lock<std::mutex> lk(e->mu);
    e->cv.wait(lk, [&] { return e->completed >= ticket; });
  });
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
  s->push([dst, src, n] {
    fakecuda::random_delay();
    memcpy(dst, src, n);   // the source is read now, not when the copy was enqueued
  });
  return cudaSuccess;
}
