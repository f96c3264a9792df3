// Fused worker step for sm_100a: whole training steps of the 784-H-10 MLP inside ONE persistent kernel.
//
// The reference's hot loop is `sess.run([train_op, loss, global_step], feed_dict)`
// (/root/reference/distributed_server-basic.py:110-113): fetch the variables from the ps (DS:41-47 placed by
// DS:88-89), forward (DS:49-51), loss (DS:52-53), backward + send the gradients to the ps (DS:103). Here that whole
// iteration is one pass of one 8-CTA thread-block cluster, repeated for as many steps as the launch was given —
// no kernel boundary, no host round trip and no global-memory hand-off between the phases of a step:
//
//   pull + forward   CTA r of the cluster owns input features [32 kb_r, 32 ke_r): warp 0 TMA-loads that K-slice of
//                    W *straight out of the owning ps shard's HBM* (NVLink peer mapping) and of x into 128B-swizzled
//                    shared memory; warp 1 issues tcgen05.mma (tf32, fp32 accumulators in TMEM): a split-K partial
//                    of pre^T [128 hidden x 32 batch].
//   reduce-scatter   the partial accumulators are read back with tcgen05.ld and scattered through distributed
//                    shared memory (st.shared::cluster.v4): CTA r ends up with batch rows 4r..4r+3 for all hidden
//                    units, adds the bias (peer load) and applies the ReLU.
//   head             logits / softmax / loss (book or xent) / accuracy / dlogits for those 4 rows, then
//                    dpre = (dlogits . W_last) * relu' and the partial dW_last / db_hidden / db_last sums.
//   all-gather       dpre^T is the A operand of the dW GEMM: every CTA writes its 16-byte chunk of each 128-byte
//                    row (K-major, SWIZZLE_128B) into the shared memory of all 8 CTAs; the small-gradient partials
//                    are reduce-scattered the same way (16 hidden units per CTA).
//   dW + push        CTA r computes dW[:, its K-slice] = dpre^T (128 x 32) . x (32 x slice) with tcgen05.mma (x is
//                    still in shared memory, MN-major view) and its epilogue *is the gradient push*: TMEM -> registers
//                    -> st.global.v4 into this worker's mailbox slot in the ps shard's HBM (or red.add into the master
//                    copy for async SGD), one cumulative st.release flag per tile.
//
// Steps are claimed dynamically (atomic counter), so `lanes` = gridDim.y clusters of one launch work on different
// steps concurrently and a cluster that is scheduled late simply takes fewer steps. Within a cluster the next
// step's W / x loads are issued as soon as the forward MMAs of the current step have drained the buffers
// (`strict` = 0), i.e. the pull of step s+1 overlaps the backward half of step s.
//
// SURVEY K1-K6, K12, X3-X5 for the flagship model; deeper / wider / bf16 models use the per-layer kernels
// (gemm_sm100.cu, head_sm100.cu) chained in a CUDA graph.
#include "common.cuh"
#include "fused.h"

namespace dm {

constexpr int kFsThreads = 192;
constexpr int kFsRows = 4;             // batch rows per CTA (32 / kFusedCluster)
constexpr int kFsMaxC = 16;            // classes, padded
constexpr int kFsPart = 12;            // floats per (source CTA, hidden unit): dW_last[0..10], db_hidden
constexpr int kFsWlStride = 132;       // smem row stride of W_last: 16 class rows land in distinct bank groups
constexpr uint32_t kFsTmemCols = 256;  // fwd accumulator: columns [0, 32); dW accumulator: columns [128, 256)
constexpr uint32_t kFsAcc2Col = 128;
constexpr uint32_t kFsWChunk = 128 * 128;   // one k-chunk of W: 128 rows x 128 B
constexpr uint32_t kFsXChunk = 32 * 128;    // one k-chunk of x: 32 rows x 128 B

// shared-memory carve-up (byte offsets from the 1024-aligned base; identical in every CTA of the cluster)
constexpr uint32_t kOffW = 0;                                        // 4 x 16 KB   W k-chunks (A of the forward)
constexpr uint32_t kOffXk = kOffW + kFusedMaxChunks * kFsWChunk;     // 4 x 4 KB    x k-chunks (B of the forward)
constexpr uint32_t kOffXmn = kOffXk + kFusedMaxChunks * kFsXChunk;   // 4 x 4 KB    x slabs    (B of the dW GEMM)
constexpr uint32_t kOffA = kOffXmn + kFusedMaxChunks * kFsXChunk;    // 16 KB       dpre^T     (A of the dW GEMM)
constexpr uint32_t kOffRed = kOffA + 128 * 128;                      // [8][128][4] reduce-scatter landing zone
constexpr uint32_t kOffSg = kOffRed + 8 * 128 * 4 * 4;               // [8][16][12] small-gradient landing zone
constexpr uint32_t kOffVal = kOffSg + 8 * 16 * kFsPart * 4;          // [4][128]    activations of my rows
constexpr uint32_t kOffWl = kOffVal + kFsRows * 128 * 4;             // [16][132]   W_last
constexpr uint32_t kOffDl = kOffWl + kFsMaxC * kFsWlStride * 4;      // [4][16]     dlogits of my rows
constexpr uint32_t kOffDbl = kOffDl + kFsRows * kFsMaxC * 4;         // [8][16]     db_last partials (CTA 0)
constexpr uint32_t kOffScal = kOffDbl + 8 * kFsMaxC * 4;             // [8][2]      loss / correct partials (CTA 0)
constexpr uint32_t kOffRedw = kOffScal + 8 * 2 * 4;                  // [8]         warp partials
constexpr uint32_t kOffCtl = kOffRedw + 8 * 4;                       // next-step words [2]
constexpr uint32_t kOffBar = kOffCtl + 32;                           // 4 mbarriers + TMEM slot
constexpr uint32_t kFsSmemUsed = kOffBar + 64;
constexpr uint32_t kFsSmemBytes = kFsSmemUsed + 1024;                // + alignment slack
static_assert(kOffA % 1024 == 0 && kOffXk % 1024 == 0 && kOffXmn % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(kOffBar % 8 == 0 && kOffRed % 16 == 0 && kOffSg % 16 == 0, "alignment");
static_assert(kFsSmemBytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void cluster_sync_all() {
  cluster_barrier_arrive_release();
  cluster_barrier_wait_acquire();
}

struct FsPush {
  float* base;
  uint32_t* flags;
};
__device__ __forceinline__ FsPush fs_resolve(const PushTarget& t, uint32_t seq) {
  FsPush r{t.base, t.flags};
  if (t.mode == PUSH_MAILBOX) {
    const uint32_t slot = seq % t.nslots;
    r.base = t.base + static_cast<uint64_t>(slot) * t.slot_stride;
    r.flags = t.flags + static_cast<uint64_t>(slot) * t.flag_slot_stride;
  }
  return r;
}
__device__ __forceinline__ void fs_push_value(const PushTarget& t, float* dst, float v) {
  if (t.mode == PUSH_ATOMIC) red_add_sys_f32(dst, t.scale * v);
  else *dst = v;
}

__device__ __forceinline__ uint64_t fs_row0(const FusedParams& p, uint32_t step) {
  return (p.row_start + static_cast<uint64_t>(step) * p.row_stride) % p.row_wrap;
}

// spin until every shard has acknowledged push `need` of this worker (acks are written into local memory by the ps)
__device__ __forceinline__ void fs_wait_acks(const FusedParams& p, uint32_t need, const char* who) {
  const uint64_t t0 = globaltimer_ns();
  for (int k = 0; k < p.n_shards; ++k) {
    const volatile uint32_t* ib = reinterpret_cast<const volatile uint32_t*>(p.shard[k].inbox);
    if (ib == nullptr) continue;
    while (static_cast<int32_t>(ib[0] - need) < 0) {
      if (globaltimer_ns() - t0 > DM_SPIN_TIMEOUT_NS) {
        printf("[dm] fused step (%s): ps shard %d ack timeout (need=%u ack=%u)\n", who, k, need, ib[0]);
        __trap();
      }
    }
  }
}

__global__ void __launch_bounds__(kFsThreads, 1)
fused_step_kernel(const __grid_constant__ FusedMaps maps, const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const FusedSlice sl = p.slice[rank];
  const int nkc = sl.kc_count;

  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + kOffBar);   // W + x (K-major) of a step have landed
  uint64_t* bar_xmn = bar_w + 1;                                    // x (MN-major) of a step has landed
  uint64_t* bar_acc1 = bar_w + 2;                                   // forward MMAs complete
  uint64_t* bar_acc2 = bar_w + 3;                                   // dW MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
  volatile uint32_t* ctl_next = reinterpret_cast<volatile uint32_t*>(smem + kOffCtl);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&maps.w[rank]);
    prefetch_tensormap(&maps.xk);
    prefetch_tensormap(&maps.xmn);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_xmn, 1);
      mbar_init(bar_acc1, 1);
      mbar_init(bar_acc2, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kFsTmemCols);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // Distributed shared memory of a peer CTA may only be touched once that CTA is known to be running.
  cluster_sync_all();
  // ---- first step of this cluster: claimed by CTA 0, broadcast to everybody ----
  auto claim = [&]() -> uint32_t {
    if (*reinterpret_cast<volatile uint32_t*>(p.stop_word) != 0u) return kFusedNoStep;
    const uint32_t g = atomicAdd(p.step_counter, 1u);
    return g < p.n_steps ? g : kFusedNoStep;
  };
  auto broadcast_next = [&](uint32_t which, uint32_t g) {
#pragma unroll
    for (uint32_t d = 0; d < kFusedCluster; ++d)
      st_shared_cluster_u32(mapa_shared_cluster(sbase + kOffCtl + which * 4u, d), g);
  };
  const bool is_claimer = rank == 0 && threadIdx.x == 64;   // epilogue thread et == 0 of CTA 0
  if (is_claimer) broadcast_next(0, claim());
  cluster_sync_all();
  uint32_t cur = ctl_next[0];

  const bool dbg = p.debug_ts != nullptr && blockIdx.y == 0 && rank == 0;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    auto issue_wx = [&](uint32_t step) {
      if (nkc == 0) { mbar_arrive(bar_w); return; }
      const int row0 = static_cast<int>(fs_row0(p, step));
      mbar_arrive_expect_tx(bar_w, static_cast<uint32_t>(nkc) * (kFsWChunk + kFsXChunk));
      for (int i = 0; i < nkc; ++i) {
        const int k0 = (sl.kc_begin + i) * 32;
        tma_load_2d(smem + kOffW + i * kFsWChunk, &maps.w[rank], bar_w, k0, 0);      // peer HBM -> smem (the pull)
        tma_load_2d(smem + kOffXk + i * kFsXChunk, &maps.xk, bar_w, k0, row0);
      }
    };
    auto issue_xmn = [&](uint32_t step) {
      if (nkc == 0) { mbar_arrive(bar_xmn); return; }
      const int row0 = static_cast<int>(fs_row0(p, step));
      mbar_arrive_expect_tx(bar_xmn, static_cast<uint32_t>(nkc) * kFsXChunk);
      for (int i = 0; i < nkc; ++i)
        tma_load_2d(smem + kOffXmn + i * kFsXChunk, &maps.xmn, bar_xmn, (sl.kc_begin + i) * 32, row0);
    };
    if (lane == 0 && cur != kFusedNoStep) {
      issue_wx(cur);
      issue_xmn(cur);
    }
    __syncwarp();
    for (uint32_t jl = 0; cur != kFusedNoStep; ++jl) {
      cluster_sync_all();                                   // #1: forward accumulators are out of TMEM / smem
      const uint32_t nxt = ctl_next[(jl + 1) & 1];
      if (lane == 0 && !p.strict && nxt != kFusedNoStep) issue_wx(nxt);   // W / x buffers are free: pull ahead
      __syncwarp();
      cluster_sync_all();                                   // #2
      cluster_sync_all();                                   // #3: dW MMAs of this step are complete
      if (lane == 0 && nxt != kFusedNoStep) {
        if (p.strict) {
          // reference-exact read-your-writes: the next pull starts only after this step's push is applied
          if (p.shard[0].push.mode == PUSH_MAILBOX) fs_wait_acks(p, p.seq_base + cur + 1u, "strict pull");
          issue_wx(nxt);
        }
        issue_xmn(nxt);
      }
      __syncwarp();
      cur = nxt;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc_f = make_idesc(MmaKind<float>::kFormat, false, false, 128, 32);
    const uint32_t idesc_w = make_idesc(MmaKind<float>::kFormat, false, true, 128, 32u * static_cast<uint32_t>(nkc > 0 ? nkc : 1));
    for (uint32_t jl = 0; cur != kFusedNoStep; ++jl) {
      mbar_wait(bar_w, jl & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        for (int i = 0; i < nkc; ++i) {
          const uint32_t a_addr = sbase + kOffW + i * kFsWChunk;
          const uint32_t b_addr = sbase + kOffXk + i * kFsXChunk;
#pragma unroll
          for (int j = 0; j < 4; ++j)   // K-major: advance 32 B inside the 128 B swizzle atom; SBO = 8 rows x 128 B
            MmaKind<float>::mma(tmem_base, make_smem_desc_sw128(a_addr + j * 32, 16, 1024),
                                make_smem_desc_sw128(b_addr + j * 32, 16, 1024), idesc_f, (i > 0 || j > 0) ? 1u : 0u);
        }
        tcgen05_commit(bar_acc1);
      }
      __syncwarp();
      cluster_sync_all();                                   // #1
      tcgen05_fence_before();
      cluster_sync_all();                                   // #2: dpre^T (A operand) is complete in my smem
      tcgen05_fence_after();
      fence_proxy_async();                                  // generic-proxy (DSMEM) writes -> tensor-core reads
      mbar_wait(bar_xmn, jl & 1);
      tcgen05_fence_after();
      if (lane == 0) {
        if (nkc > 0) {
          const uint32_t a_addr = sbase + kOffA;
          const uint32_t b_addr = sbase + kOffXmn;
#pragma unroll
          for (int j = 0; j < 4; ++j)   // k = batch: 4 steps of 8 rows. A K-major (32 B per step); B MN-major tf32:
                                        // 8 k-rows x 128 B per step, LBO = slab stride, SBO = 4-row group stride
            MmaKind<float>::mma(tmem_base + kFsAcc2Col, make_smem_desc_sw128(a_addr + j * 32, 16, 1024),
                                make_smem_desc_sw128(b_addr + j * (8 * 128), kFsXChunk, 512, 1), idesc_w, j > 0 ? 1u : 0u);
        }
        tcgen05_commit(bar_acc2);
      }
      __syncwarp();
      cluster_sync_all();                                   // #3
      cur = ctl_next[(jl + 1) & 1];
    }
  } else {
    // =========================== epilogue / head warps ===========================
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int et = q * 32 + lane;                 // TMEM lane == hidden unit == tile row
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int B = p.B, H = p.H, C = p.C;
    const bool unit_ok = et < H;
    float* red = reinterpret_cast<float*>(smem + kOffRed);
    float* sg = reinterpret_cast<float*>(smem + kOffSg);
    float* val = reinterpret_cast<float*>(smem + kOffVal);
    float* wls = reinterpret_cast<float*>(smem + kOffWl);
    float* dls = reinterpret_cast<float*>(smem + kOffDl);
    float* dbl = reinterpret_cast<float*>(smem + kOffDbl);
    float* scal = reinterpret_cast<float*>(smem + kOffScal);
    float* redw = reinterpret_cast<float*>(smem + kOffRedw);
    const bool mailbox = p.shard[0].push.mode == PUSH_MAILBOX;
    const uint32_t scope = p.shard[0].push.gpu_scope;

    for (uint32_t jl = 0; cur != kFusedNoStep; ++jl) {
      const uint32_t seq = p.seq_base + cur + 1u;
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 0] = clock64();
      // ---- step prologue: everything that does not depend on the forward result is requested now ----
      if (is_claimer) broadcast_next((jl + 1) & 1, claim());
      const float bias = unit_ok ? p.bias_h[et] : 0.f;          // peer loads from the ps shard(s)
      float wl[kFsMaxC];
#pragma unroll
      for (int c = 0; c < kFsMaxC; ++c) wl[c] = (unit_ok && c < C) ? p.w_last[static_cast<size_t>(c) * H + et] : 0.f;
      // head role of this thread (threads et < 64): (row rr, class hc)
      const int rr = (et >> 4) & 3, hc = et & 15;
      const int brow = static_cast<int>(rank) * kFsRows + rr;
      const bool hrow_ok = et < 64 && brow < B;
      const bool hc_ok = hc < C;
      float ylab = 0.f, blast = 0.f;
      if (et < 64) {
        if (hrow_ok && hc_ok) ylab = p.y_base[(fs_row0(p, cur) + static_cast<uint64_t>(brow)) * C + hc];
        if (hc_ok) blast = p.b_last[hc];
      }
      if (et == 127 && mailbox && seq > p.nslots) fs_wait_acks(p, seq - p.nslots, "flow control");

      // ---- forward partial -> reduce-scatter over the cluster ----
      mbar_wait(bar_acc1, jl & 1);
      tcgen05_fence_after();
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 1] = clock64();
      {
        uint32_t r[32];
        if (nkc > 0) {
          tmem_ld_32x32b_x32_nowait(taddr, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const uint32_t dst_off = kOffRed + ((rank * 128u + static_cast<uint32_t>(et)) << 4);
#pragma unroll
        for (uint32_t d = 0; d < kFusedCluster; ++d)
          st_shared_cluster_v4f32(mapa_shared_cluster(sbase + dst_off, d), __uint_as_float(r[4 * d]),
                                  __uint_as_float(r[4 * d + 1]), __uint_as_float(r[4 * d + 2]),
                                  __uint_as_float(r[4 * d + 3]));
      }
      tcgen05_fence_before();
      cluster_sync_all();                                   // #1
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 2] = clock64();
      float hv[kFsRows];
      {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kFusedCluster; ++s) {
          const float4 t = *reinterpret_cast<const float4*>(red + (s * 128 + et) * 4);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        hv[0] = unit_ok ? fmaxf(acc.x + bias, 0.f) : 0.f;
        hv[1] = unit_ok ? fmaxf(acc.y + bias, 0.f) : 0.f;
        hv[2] = unit_ok ? fmaxf(acc.z + bias, 0.f) : 0.f;
        hv[3] = unit_ok ? fmaxf(acc.w + bias, 0.f) : 0.f;
      }
      // ---- head on my 4 batch rows ----
#pragma unroll
      for (int i = 0; i < kFsRows; ++i) val[i * 128 + et] = hv[i];
#pragma unroll
      for (int c = 0; c < kFsMaxC; ++c) wls[c * kFsWlStride + et] = wl[c];
      named_bar_sync(1, 128);
      if (et < 64) {   // warps with q == 0, 1 (warp-uniform)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* hrow = val + rr * 128;
        const float* wrow = wls + hc * kFsWlStride;
#pragma unroll 8
        for (int k = 0; k < 128; k += 4) {   // units >= H are zero in both operands
          const float4 hvv = *reinterpret_cast<const float4*>(hrow + k);
          const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
          a0 = fmaf(hvv.x, wv.x, a0); a1 = fmaf(hvv.y, wv.y, a1);
          a2 = fmaf(hvv.z, wv.z, a2); a3 = fmaf(hvv.w, wv.w, a3);
        }
        const float y = ylab;
        const float z = hc_ok ? (a0 + a1) + (a2 + a3) + blast : -INFINITY;
        float zmax = z, ybest = hc_ok ? y : -INFINITY, ysum = y;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
          ybest = fmaxf(ybest, __shfl_xor_sync(0xffffffffu, ybest, o));
          ysum += __shfl_xor_sync(0xffffffffu, ysum, o);
        }
        int zarg = (hc_ok && z == zmax) ? hc : kFsMaxC, yarg = (hc_ok && y == ybest) ? hc : kFsMaxC;  // first maximum wins
        const float e = hc_ok ? __expf(z - zmax) : 0.f;
        float esum = e;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          zarg = min(zarg, __shfl_xor_sync(0xffffffffu, zarg, o));
          yarg = min(yarg, __shfl_xor_sync(0xffffffffu, yarg, o));
          esum += __shfl_xor_sync(0xffffffffu, esum, o);
        }
        const float pc = e / esum;
        float dl = 0.f, lc = 0.f;
        if (p.loss_kind == LOSS_BOOK) {
          // L = -(1/(B*C)) sum y*log(clip(p,1e-10,1));  dL/dp = -y/(B*C*p) where the clip passes gradient (DS:52-53)
          const float k = 1.f / (static_cast<float>(B) * static_cast<float>(C));
          const bool pass = hc_ok && (pc >= 1e-10f) && (pc <= 1.0f);
          const float g = pass ? (-k * y / pc) : 0.f;
          float gp = g * pc;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) gp += __shfl_xor_sync(0xffffffffu, gp, o);
          dl = pc * (g - gp);
          lc = hc_ok ? -k * y * __logf(fminf(fmaxf(pc, 1e-10f), 1.0f)) : 0.f;
        } else {
          // L = (1/B) sum_b -sum_c y*log_softmax(z);  dz = (p*sum(y) - y)/B   (DS:35)
          const float k = 1.f / static_cast<float>(B);
          dl = k * (pc * ysum - y);
          lc = hc_ok ? -k * y * (z - (zmax + __logf(esum))) : 0.f;
        }
        dls[rr * kFsMaxC + hc] = (hrow_ok && hc_ok) ? dl : 0.f;
        float loss_part = hrow_ok ? lc : 0.f;
        float corr_part = (hrow_ok && hc == 0 && zarg == yarg) ? 1.f : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
          corr_part += __shfl_xor_sync(0xffffffffu, corr_part, o);
        }
        if (lane == 0) { redw[q] = loss_part; redw[4 + q] = corr_part; }
      }
      named_bar_sync(1, 128);
      // ---- gradients of my rows; thread -> hidden unit et ----
      float dw[kFsPart];
#pragma unroll
      for (int c = 0; c < kFsPart; ++c) dw[c] = 0.f;
      float dp[kFsRows];
      float dbh = 0.f;
#pragma unroll
      for (int i = 0; i < kFsRows; ++i) {
        float dh = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < kFsMaxC / 4; ++c4) {
          const float4 d = *reinterpret_cast<const float4*>(dls + i * kFsMaxC + 4 * c4);   // broadcast reads
          dh = fmaf(d.x, wl[4 * c4 + 0], dh); dh = fmaf(d.y, wl[4 * c4 + 1], dh);
          dh = fmaf(d.z, wl[4 * c4 + 2], dh); dh = fmaf(d.w, wl[4 * c4 + 3], dh);
          if (4 * c4 + 0 < kFsPart - 1) dw[4 * c4 + 0] = fmaf(d.x, hv[i], dw[4 * c4 + 0]);
          if (4 * c4 + 1 < kFsPart - 1) dw[4 * c4 + 1] = fmaf(d.y, hv[i], dw[4 * c4 + 1]);
          if (4 * c4 + 2 < kFsPart - 1) dw[4 * c4 + 2] = fmaf(d.z, hv[i], dw[4 * c4 + 2]);
          if (4 * c4 + 3 < kFsPart - 1) dw[4 * c4 + 3] = fmaf(d.w, hv[i], dw[4 * c4 + 3]);
        }
        dp[i] = hv[i] > 0.f ? dh : 0.f;       // rows >= B carry dlogits == 0 -> dp == 0
        dbh += dp[i];
      }
      dw[kFsPart - 1] = dbh;
      {
        // all-gather of dpre^T: row et (128 B = 32 batch values), my 16-byte chunk is number `rank`; K-major
        // SWIZZLE_128B: chunk j of row r sits at r * 128 + ((j ^ (r & 7)) << 4)
        const uint32_t a_off = kOffA + static_cast<uint32_t>(et) * 128u + ((rank ^ (static_cast<uint32_t>(et) & 7u)) << 4);
#pragma unroll
        for (uint32_t d = 0; d < kFusedCluster; ++d)
          st_shared_cluster_v4f32(mapa_shared_cluster(sbase + a_off, d), dp[0], dp[1], dp[2], dp[3]);
        // reduce-scatter of the small-gradient partials: hidden units [16 d, 16 d + 16) are summed by CTA d
        const uint32_t sg_off = kOffSg + ((rank * 16u + (static_cast<uint32_t>(et) & 15u)) * kFsPart) * 4u;
        const uint32_t sg_dst = mapa_shared_cluster(sbase + sg_off, static_cast<uint32_t>(et) >> 4);
        st_shared_cluster_v4f32(sg_dst, dw[0], dw[1], dw[2], dw[3]);
        st_shared_cluster_v4f32(sg_dst + 16u, dw[4], dw[5], dw[6], dw[7]);
        st_shared_cluster_v4f32(sg_dst + 32u, dw[8], dw[9], dw[10], dw[11]);
        if (et < kFsMaxC) {
          float sdl = 0.f;
#pragma unroll
          for (int i = 0; i < kFsRows; ++i) sdl += dls[i * kFsMaxC + et];
          st_shared_cluster_f32(mapa_shared_cluster(sbase + kOffDbl + (rank * kFsMaxC + et) * 4u, 0), sdl);
        }
        if (et == 0) {
          st_shared_cluster_f32(mapa_shared_cluster(sbase + kOffScal + (rank * 2u) * 4u, 0), redw[0] + redw[1]);
          st_shared_cluster_f32(mapa_shared_cluster(sbase + kOffScal + (rank * 2u + 1u) * 4u, 0), redw[4] + redw[5]);
        }
      }
      fence_proxy_async();
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 3] = clock64();
      cluster_sync_all();                                   // #2
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 4] = clock64();

      // ---- small gradients: sum the 8 sources for my 16 hidden units and push them ----
      const FsPush pw_l = fs_resolve(p.shard[p.shard_wl].push, seq);
      const FsPush pb_h = fs_resolve(p.shard[p.shard_bh].push, seq);
      const FsPush pb_l = fs_resolve(p.shard[p.shard_bl].push, seq);
      {
        const int u = et & 15, vg = et >> 4;   // values vg (0..7) and vg + 8 (8..11) of unit u
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int s = 0; s < kFusedCluster; ++s) {
          s0 += sg[(s * 16 + u) * kFsPart + vg];
          if (vg < 4) s1 += sg[(s * 16 + u) * kFsPart + vg + 8];
        }
        const int unit = static_cast<int>(rank) * 16 + u;
        if (unit < H) {
          if (vg < C) fs_push_value(p.shard[p.shard_wl].push, pw_l.base + p.off_wl + static_cast<size_t>(vg) * H + unit, s0);
          if (vg < 3 && vg + 8 < C)
            fs_push_value(p.shard[p.shard_wl].push, pw_l.base + p.off_wl + static_cast<size_t>(vg + 8) * H + unit, s1);
          if (vg == 3) fs_push_value(p.shard[p.shard_bh].push, pb_h.base + p.off_bh + unit, s1);
        }
      }
      float loss_tot = 0.f, corr_tot = 0.f;
      if (rank == 0) {
        if (et < C) {
          float s = 0.f;
#pragma unroll
          for (int src = 0; src < kFusedCluster; ++src) s += dbl[src * kFsMaxC + et];
          fs_push_value(p.shard[p.shard_bl].push, pb_l.base + p.off_bl + et, s);
        }
        if (et == 0) {
#pragma unroll
          for (int src = 0; src < kFusedCluster; ++src) { loss_tot += scal[src * 2]; corr_tot += scal[src * 2 + 1]; }
        }
      }

      // ---- dW tile of my K-slice: the epilogue is the gradient push ----
      mbar_wait(bar_acc2, jl & 1);
      tcgen05_fence_after();
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 5] = clock64();
      const PushTarget& tw = p.shard[sl.shard].push;
      const FsPush pw = fs_resolve(tw, seq);
      if (nkc > 0) {
        float* row = pw.base + sl.w_offset + static_cast<size_t>(et) * p.ldw + static_cast<size_t>(sl.kc_begin) * 32;
        const int col0 = sl.kc_begin * 32;
        const float sc = tw.scale;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32_nowait(taddr + kFsAcc2Col, ra);
#pragma unroll
        for (int i = 0; i < kFusedMaxChunks; ++i) {
          if (i >= nkc) break;
          tmem_ld_wait();
          uint32_t (&cur_r)[32] = (i & 1) ? rb : ra;
          uint32_t (&nxt_r)[32] = (i & 1) ? ra : rb;
          if (i + 1 < nkc) tmem_ld_32x32b_x32_nowait(taddr + kFsAcc2Col + 32u * (i + 1), nxt_r);   // overlaps the stores
          if (unit_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int c = col0 + i * 32 + j;
              if (c < p.I) {   // I is a multiple of 4 (checked on the host): whole 16-byte groups
                float* dst = row + i * 32 + j;
                const float v0 = __uint_as_float(cur_r[j]), v1 = __uint_as_float(cur_r[j + 1]);
                const float v2 = __uint_as_float(cur_r[j + 2]), v3 = __uint_as_float(cur_r[j + 3]);
                if (tw.mode == PUSH_ATOMIC) red_add_sys_v4f32(dst, sc * v0, sc * v1, sc * v2, sc * v3);
                else st_global_v4f32(dst, v0, v1, v2, v3);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      // Publish: the barrier orders every lane's P2P stores before the single cumulative release of the tile flag.
      named_bar_sync(1, 128);
      if (mailbox && et == 0 && nkc > 0) st_release_scoped_u32(pw.flags + sl.flag_index, seq, tw.gpu_scope);
      if (dbg && et == 0 && jl < 8) p.debug_ts[jl * 8 + 6] = clock64();
      cluster_sync_all();                                   // #3: every CTA's small-gradient stores are ordered before
      if (rank == 0 && et == 0) {                           //     the flags CTA 0 publishes now
        uint32_t gstep = seq;
        if (mailbox) {
          fence_acq_rel_scoped(scope);
          st_relaxed_sys_u32(pw_l.flags + p.flag_wl, seq);
          st_relaxed_sys_u32(pb_h.flags + p.flag_bh, seq);
          st_relaxed_sys_u32(pb_l.flags + p.flag_bl, seq);
          const volatile uint32_t* ib = reinterpret_cast<const volatile uint32_t*>(p.shard[0].inbox);
          if (ib != nullptr) {
            // global_step as of our last acknowledged push + our own pushes since then: exact with one worker,
            // a lower bound under concurrency (the reference's fetched value is equally unordered w.r.t. peers)
            const uint32_t ack = ib[0];
            __threadfence();
            gstep = ib[1] + (seq - ack);
          }
        } else if (p.shard[0].push.mode == PUSH_ATOMIC && p.ps_global_step != nullptr) {
          gstep = atom_add_sys_u32(p.ps_global_step, 1u) + 1u;   // async SGD: this push *is* global step gstep
        }
        StepResult res;
        res.loss = loss_tot;
        res.global_step = gstep;
        res.correct = static_cast<uint32_t>(corr_tot + 0.5f);
        res.seq = seq;
        p.results[cur] = res;                               // pinned host memory: a 16-byte posted write
        atomicAdd(p.seq_word, 1u);
        if (p.stop_at != 0u && gstep >= p.stop_at) *reinterpret_cast<volatile uint32_t*>(p.stop_word) = 1u;
        if (dbg && jl < 8) p.debug_ts[jl * 8 + 7] = clock64();
      }
      cur = ctl_next[(jl + 1) & 1];
    }
  }

  // ---- teardown ----
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kFsTmemCols);
  cluster_sync_all();   // nobody leaves while a peer might still address its shared memory
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
cudaError_t prepare_fused_kernel() {
  cudaError_t e = cudaFuncSetAttribute(fused_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kFsSmemBytes));
  if (e != cudaSuccess) return e;
  cudaFuncAttributes a;
  return cudaFuncGetAttributes(&a, fused_step_kernel);   // force-load (see preload_ps_kernels)
}

size_t fused_smem_bytes() { return kFsSmemBytes; }

cudaError_t launch_fused_step(const FusedMaps& maps, const FusedParams& p, int lanes, cudaStream_t stream) {
  if (lanes < 1 || p.B < 1 || p.B > 32 || p.H < 1 || p.H > 128 || p.C < 1 || p.C > kFsPart - 1 || (p.I & 3) ||
      p.n_shards < 1 || p.n_shards > kFusedMaxShards || p.nslots < 1)
    return cudaErrorInvalidValue;
  for (int r = 0; r < kFusedCluster; ++r)
    if (p.slice[r].kc_count < 0 || p.slice[r].kc_count > kFusedMaxChunks) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kFusedCluster, lanes, 1);
  cfg.blockDim = dim3(kFsThreads);
  cfg.dynamicSmemBytes = kFsSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kFusedCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fused_step_kernel, maps, p);
}

// How many clusters of the fused kernel can be co-resident on the device (upper bound for `lanes`).
cudaError_t fused_max_lanes(int* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kFusedCluster, 1, 1);
  cfg.blockDim = dim3(kFsThreads);
  cfg.dynamicSmemBytes = kFsSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kFusedCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaOccupancyMaxActiveClusters(out, fused_step_kernel, &cfg);
}

}  // namespace dm
