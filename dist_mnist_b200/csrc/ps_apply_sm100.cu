// Parameter-server shard kernels for sm_100a.
//
// ps_serve_kernel is the PS "process" of the reference (tf.train.Server(...).join(),
// /root/reference/distributed_server-basic.py:80-83) turned into a persistent GPU kernel: it owns the
// shard (params + Adam slots + global_step, DS:88-91,102-103), polls the per-item flags that workers
// publish with st.release.sys after their in-kernel P2P gradient stores, reduces whatever is ready
// (many-to-one) and applies the optimizer in the same pass (ApplyAdam / SGD), bumps
// global_step (K8) and acknowledges into the worker's inbox over NVLink. No host, NCCL or gRPC in the loop.
//
// Work decomposition: items (2-D blocks of the arena) are dealt round-robin to CTAs; a CTA never shares
// an item, so item state (Adam step count / beta powers) needs no synchronisation. Pushes of one worker
// are consumed in order per item; different workers interleave arbitrarily (Hogwild, like the reference's
// unlocked async apply).
#include "common.cuh"
#include "protocol.h"

namespace dm {

constexpr int kPsThreads = 256;

struct UpdateCtx {
  int opt;
  float lr, beta1, beta2, eps;
};

__device__ __forceinline__ void adam_step(float& p, float& m, float& v, float g, float lr_t, float beta1,
                                          float beta2, float eps) {
  // TF1 AdamOptimizer: m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; p <- p - lr_t m / (sqrt(v) + eps)
  m = fmaf(beta1, m, (1.f - beta1) * g);
  v = fmaf(beta2, v, (1.f - beta2) * g * g);
  p -= lr_t * m / (sqrtf(v) + eps);
}

// Apply the gradients of the ready workers (bitmask) to one item. Called by the whole CTA.
__device__ void apply_item(const PsServeParams& P, const PsItem it, PsItemState st, uint32_t mask,
                           const uint32_t* s_seq) {
  const int tid = threadIdx.x;
  const uint64_t wstride = static_cast<uint64_t>(P.nslots) * P.arena_elems;
  // per-push schedule of lr_t (same for every element of the item)
  const int total = it.rows * it.cols;
  const bool vec = ((it.cols & 3) == 0) && ((it.ld & 3) == 0) && ((it.offset & 3) == 0);
  const int step = vec ? 4 : 1;
  for (int e = tid * step; e < total; e += kPsThreads * step) {
    const int r = e / it.cols;
    const int c = e - r * it.cols;
    const uint64_t a = it.offset + static_cast<uint64_t>(r) * it.ld + c;
    float pv[4], mv[4] = {0, 0, 0, 0}, vv[4] = {0, 0, 0, 0};
    if (vec) {
      const float4 t = *reinterpret_cast<const float4*>(P.params + a);
      pv[0] = t.x; pv[1] = t.y; pv[2] = t.z; pv[3] = t.w;
      if (P.opt == OPT_ADAM) {
        const float4 tm = *reinterpret_cast<const float4*>(P.adam_m + a);
        const float4 tv = *reinterpret_cast<const float4*>(P.adam_v + a);
        mv[0] = tm.x; mv[1] = tm.y; mv[2] = tm.z; mv[3] = tm.w;
        vv[0] = tv.x; vv[1] = tv.y; vv[2] = tv.z; vv[3] = tv.w;
      }
    } else {
      pv[0] = P.params[a];
      if (P.opt == OPT_ADAM) { mv[0] = P.adam_m[a]; vv[0] = P.adam_v[a]; }
    }
    float b1p = st.beta1_pow, b2p = st.beta2_pow;
    float gsum[4] = {0, 0, 0, 0};
    uint32_t mm = mask;
    while (mm) {
      const int w = __ffs(mm) - 1;
      mm &= mm - 1;
      const uint32_t slot = s_seq[w] % P.nslots;
      const float* gsrc = P.mailbox + static_cast<uint64_t>(w) * wstride + static_cast<uint64_t>(slot) * P.arena_elems + a;
      float g[4];
      if (vec) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(gsrc));  // L2 only: written remotely over NVLink
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
      } else {
        g[0] = __ldcg(gsrc);
      }
      if (P.apply_mode == APPLY_MERGED) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gsum[j] += g[j];
        continue;
      }
      if (P.opt == OPT_ADAM) {
        b1p *= P.beta1;
        b2p *= P.beta2;
        const float lr_t = P.lr * sqrtf(1.f - b2p) / (1.f - b1p);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < step) adam_step(pv[j], mv[j], vv[j], g[j], lr_t, P.beta1, P.beta2, P.eps);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < step) pv[j] = fmaf(-P.lr, g[j], pv[j]);
      }
    }
    if (P.apply_mode == APPLY_MERGED) {
      if (P.opt == OPT_ADAM) {
        b1p *= P.beta1;
        b2p *= P.beta2;
        const float lr_t = P.lr * sqrtf(1.f - b2p) / (1.f - b1p);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < step) adam_step(pv[j], mv[j], vv[j], gsum[j], lr_t, P.beta1, P.beta2, P.eps);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < step) pv[j] = fmaf(-P.lr, gsum[j], pv[j]);
      }
    }
    if (vec) {
      *reinterpret_cast<float4*>(P.params + a) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      if (P.opt == OPT_ADAM) {
        *reinterpret_cast<float4*>(P.adam_m + a) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(P.adam_v + a) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
      if (P.shadow_bf16 != nullptr && (it.flags & 1)) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]);
        __nv_bfloat162 hi = __floats2bfloat162_rn(pv[2], pv[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(P.shadow_bf16 + a) = pk;
      }
    } else {
      P.params[a] = pv[0];
      if (P.opt == OPT_ADAM) { P.adam_m[a] = mv[0]; P.adam_v[a] = vv[0]; }
      if (P.shadow_bf16 != nullptr && (it.flags & 1)) {
        __nv_bfloat16 b = __float2bfloat16(pv[0]);
        P.shadow_bf16[a] = *reinterpret_cast<uint16_t*>(&b);
      }
    }
  }
}

__global__ void __launch_bounds__(kPsThreads, 1) ps_serve_kernel(const __grid_constant__ PsServeParams P) {
  __shared__ uint32_t s_mask;
  __shared__ uint32_t s_seq[kMaxWorkers];
  __shared__ uint32_t s_exit;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  uint32_t iter = 0;

  for (;;) {
    bool pending_possible = false;  // some worker may still push to one of my items
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      if (warp == 0) {
        uint32_t ready = 0, seq = 0;
        if (lane < P.n_workers) {
          seq = P.next_seq[static_cast<size_t>(lane) * P.n_items + item];
          const uint32_t slot = seq % P.nslots;
          const uint32_t f = ld_acquire_scoped_u32(
              P.flags + (static_cast<size_t>(lane) * P.nslots + slot) * P.n_items + item, P.gpu_scope);
          ready = (f == seq) ? 1u : 0u;
          s_seq[lane] = seq;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, ready);
        if (lane == 0) s_mask = mask;
      }
      __syncthreads();
      const uint32_t mask = s_mask;
      if (mask) {
        const PsItem it = P.items[item];
        const PsItemState st = P.item_state[item];
        apply_item(P, it, st, mask, s_seq);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
          const int npush = __popc(mask);
          const int nsteps = (P.apply_mode == APPLY_MERGED) ? 1 : npush;
          PsItemState ns = st;
          ns.t += nsteps;
          for (int k = 0; k < nsteps; ++k) { ns.beta1_pow *= P.beta1; ns.beta2_pow *= P.beta2; }
          P.item_state[item] = ns;
          uint32_t mm = mask;
          while (mm) {
            const int w = __ffs(mm) - 1;
            mm &= mm - 1;
            const uint32_t seq = s_seq[w];
            const uint32_t slot = seq % P.nslots;
            P.next_seq[static_cast<size_t>(w) * P.n_items + item] = seq + 1;
            const uint32_t done = atomicAdd(&P.consumed[w * P.nslots + slot], 1u) + 1u;
            if (done == static_cast<uint32_t>(P.n_items)) {
              // this worker's push `seq` is fully applied: one global step (reference DS:91,103)
              P.consumed[w * P.nslots + slot] = 0;
              const uint32_t gs = atomicAdd(P.global_step, 1u) + 1u;
              uint32_t* ib = P.inbox_table[w];
              if (ib != nullptr) {
                reinterpret_cast<volatile uint32_t*>(ib)[1] = gs;
                st_release_scoped_u32(ib, seq, P.gpu_scope);  // ack: the mailbox slot may be reused
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- exit protocol: host stop request, or every worker has left and all their pushes are applied ----
    // (checked every 32nd poll round only: the check itself costs several system-scope loads, and the poll
    //  loop's period is the PS's reaction latency)
    if (tid == 0) {
      uint32_t ex = 0;
      const bool check = (iter & 31u) == 0u;
      if (check && *P.host_stop != 0u) ex = 1;
      if (check && !ex) {
        bool all_done = true;
        for (int w = 0; w < P.n_workers && all_done; ++w) {
          const uint32_t d = ld_acquire_sys_u32(P.worker_done + w);  // = last push seq + 1, 0 while active
          if (d == 0) { all_done = false; break; }
          for (int item = blockIdx.x; item < P.n_items; item += gridDim.x)
            if (P.next_seq[static_cast<size_t>(w) * P.n_items + item] != d) { all_done = false; break; }
        }
        if (all_done) ex = 1;
      }
      s_exit = ex;
    }
    (void)pending_possible;
    __syncthreads();
    if (s_exit) break;
    ++iter;
  }
  if (tid == 0) atomicAdd(P.exit_counter, 1u);
}

cudaError_t launch_ps_serve(const PsServeParams& p, int n_ctas, cudaStream_t stream) {
  if (p.n_workers > kMaxWorkers) return cudaErrorInvalidValue;
  if (n_ctas > p.n_items) n_ctas = p.n_items;
  if (n_ctas < 1) n_ctas = 1;
  ps_serve_kernel<<<n_ctas, kPsThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// One-shot dense optimizer apply over a flat range (synchronous baselines: NCCL reduce + apply,
// host-staged gRPC stand-in on the GPU, unit tests of the update rule).
// ------------------------------------------------------------------------------------------
__global__ void dense_apply_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                   const float* __restrict__ grad, uint16_t* __restrict__ shadow, size_t n, int opt,
                                   float lr, float beta1, float beta2, float eps, float lr_t) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float p = params[i];
    const float g = grad[i];
    if (opt == OPT_ADAM) {
      float mm = m[i], vv = v[i];
      adam_step(p, mm, vv, g, lr_t, beta1, beta2, eps);
      m[i] = mm;
      v[i] = vv;
    } else {
      p = fmaf(-lr, g, p);
    }
    params[i] = p;
    if (shadow != nullptr) {
      __nv_bfloat16 b = __float2bfloat16(p);
      shadow[i] = *reinterpret_cast<uint16_t*>(&b);
    }
  }
}

cudaError_t launch_dense_apply(float* params, float* m, float* v, const float* grad, uint16_t* shadow, size_t n,
                               int opt, float lr, float beta1, float beta2, float eps, uint32_t t,
                               cudaStream_t stream) {
  float lr_t = lr;
  if (opt == OPT_ADAM) lr_t = lr * sqrtf(1.f - powf(beta2, static_cast<float>(t))) / (1.f - powf(beta1, static_cast<float>(t)));
  const int threads = 256;
  int blocks = static_cast<int>((n + threads - 1) / threads);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  dense_apply_kernel<<<blocks, threads, 0, stream>>>(params, m, v, grad, shadow, n, opt, lr, beta1, beta2, eps, lr_t);
  return cudaGetLastError();
}

// fp32 -> bf16 shadow refresh for a flat range (chief init / checkpoint restore).
__global__ void shadow_refresh_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    __nv_bfloat16 b = __float2bfloat16(src[i]);
    dst[i] = *reinterpret_cast<uint16_t*>(&b);
  }
}
cudaError_t launch_shadow_refresh(const float* src, uint16_t* dst, size_t n, cudaStream_t stream) {
  const int threads = 256;
  int blocks = static_cast<int>((n + threads - 1) / threads);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  shadow_refresh_kernel<<<blocks, threads, 0, stream>>>(src, dst, n);
  return cudaGetLastError();
}

// Stream-ordered "all my pushes are applied" fence on the worker: spins (locally — the PS writes the inbox
// into the worker's HBM) until every shard has acknowledged the worker's latest push. Lets a CUDA-event
// timed region include the PS-side apply of its last step.
__global__ void wait_ack_kernel(const uint32_t* inbox, uint32_t n_inbox, const uint32_t* seq_ptr) {
  const uint32_t seq = *reinterpret_cast<const volatile uint32_t*>(seq_ptr);
  const uint64_t t0 = globaltimer_ns();
  for (uint32_t i = 0; i < n_inbox; ++i) {
    while (static_cast<int32_t>(ld_acquire_sys_u32(inbox + 2 * i) - seq) < 0) {
      if (globaltimer_ns() - t0 > DM_SPIN_TIMEOUT_NS) {
        printf("[dm] wait_ack: shard %u stuck at ack=%u, seq=%u\n", i, inbox[2 * i], seq);
        __trap();
      }
    }
  }
}
cudaError_t launch_wait_ack(const uint32_t* inbox, uint32_t n_inbox, const uint32_t* seq_ptr, cudaStream_t stream) {
  wait_ack_kernel<<<1, 1, 0, stream>>>(inbox, n_inbox, seq_ptr);
  return cudaGetLastError();
}

// Worker leaves the session: publish last_seq + 1 into the PS's worker_done slot (peer store).
__global__ void worker_done_kernel(uint32_t* done_slot, const uint32_t* seq_ptr) {
  __threadfence_system();
  st_release_sys_u32(done_slot, *seq_ptr + 1u);
}
cudaError_t launch_worker_done(uint32_t* done_slot, const uint32_t* seq_ptr, cudaStream_t stream) {
  worker_done_kernel<<<1, 1, 0, stream>>>(done_slot, seq_ptr);
  return cudaGetLastError();
}

}  // namespace dm

namespace dm {
// Force-load every kernel of this translation unit. With CUDA's lazy module loading the *first* launch of a
// kernel may need a context-wide synchronisation; if that first launch happens while the persistent serve
// kernel is resident (and the serve kernel is waiting for the new kernel's flags) the process deadlocks.
cudaError_t preload_ps_kernels() {
  cudaFuncAttributes a;
  cudaError_t e;
  if ((e = cudaFuncGetAttributes(&a, ps_serve_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&a, dense_apply_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&a, shadow_refresh_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&a, wait_ack_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&a, worker_done_kernel)) != cudaSuccess) return e;
  return cudaSuccess;
}
}  // namespace dm
