// NVTX ranges for the native runtime (the reference has no tracing at all; the B200 plan names NVTX ranges
// next to the CUDA-event step timers and the committed ncu captures).
//
// nvtx3 is header-only: without a profiler attached a push / pop is one load of a null function pointer, so the ranges
// stay compiled in. Timeline tools (nsys, ncu --nvtx) then show the host side of a run next to the kernels:
//   dm.fexec.run / dm.fexec.chunk.plan / dm.fexec.chunk.launch / dm.fexec.harvest / dm.fexec.epoch_fill
//   dm.exec.run (graph engine), dm.loader.enable_feed
// The Python layers add theirs through dist_mnist_b200.utils.metrics.nvtx_range (torch.cuda.nvtx).
#pragma once

#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>) && !defined(DM_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
#define DM_HAVE_NVTX 1
#endif
#endif

namespace dm {

struct NvtxRange {
#ifdef DM_HAVE_NVTX
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
#else
  explicit NvtxRange(const char*) {}
#endif
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

inline void nvtx_mark(const char* name) {
#ifdef DM_HAVE_NVTX
  nvtxMarkA(name);
#else
  (void)name;
#endif
}

}  // namespace dm
