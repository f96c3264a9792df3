"""The worker task: between-graph replicated, asynchronous data-parallel training against the PS shards.

Reference parity (`/root/reference/distributed_server-basic.py`):
  * DS:87-103   each worker builds its *own* replica of the model; variables live on the ps tasks.
  * DS:108-109  `MonitoredTrainingSession(master, is_chief=(task_index == 0), ...)`: worker 0 initialises the
                variables (or restores a checkpoint), the others wait until that has happened.
  * DS:110-113  one step = pull variables, forward/backward on the worker, push gradients, PS applies Adam and
                bumps `global_step`; the worker gets `loss` and `global_step` back. No locks, no barriers.

GPU backend: a step is a PDL-linked chain of hand-written sm_100a kernels inside a CUDA graph (see `ops/`): the
forward GEMMs pull their weight tiles straight out of the PS shard's HBM with TMA over NVLink, the dW GEMM
epilogues and the classifier-head kernel push gradients into the PS mailbox (or red.add them into the master
copy for SGD) and publish per-tile flags; the persistent PS kernel applies. The native executor keeps
`cfg.lanes` steps in flight, launches `cfg.graph_steps` of them per graph, gathers batches on helper threads
and receives each step's 16-byte result in pinned host memory (csrc/executor.cu).

CPU backend: same protocol over POSIX shm with torch CPU math (BASELINE.json config 1, plumbing tests).
"""
from __future__ import annotations

import collections.abc
import ctypes as C
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _native as N
from ..cluster import ClusterSpec, Rendezvous
from ..models import mlp
from ..models.mlp import MLPSpec
from ..ops import gemm as gemm_ops
from ..ops import head as head_ops
from .config import MAX_SLOTS, EngineConfig, OptimizerConfig
from .peer_mem import Carver, Segment
from .ps import CTRL_GLOBAL_STEP, CTRL_WORKER_DONE
from .sharding import ModelLayout, VarLayout, build_layout, dw_tile_n_for


@dataclass
class StepOutput:
    loss: float
    global_step: int
    correct: int
    seq: int


_STEP_DTYPE = np.dtype([("loss", np.float32), ("global_step", np.uint32), ("correct", np.uint32), ("seq", np.uint32)])


class StepOutputs(collections.abc.Sequence):
    """Results of a native run of steps: a read-only sequence of `StepOutput` backed by the executor's result
    array (no per-step Python object is built unless the step is looked at). `.loss`, `.global_step`, `.correct`
    and `.seq` give the whole columns as numpy arrays."""

    def __init__(self, raw: np.ndarray):
        self._raw = raw

    def __len__(self) -> int:
        return int(self._raw.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return StepOutputs(self._raw[i])
        r = self._raw[i]
        return StepOutput(float(r["loss"]), int(r["global_step"]), int(r["correct"]), int(r["seq"]))

    loss = property(lambda self: self._raw["loss"])
    global_step = property(lambda self: self._raw["global_step"])
    correct = property(lambda self: self._raw["correct"])
    seq = property(lambda self: self._raw["seq"])


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class Worker:
    def __init__(self, cluster: ClusterSpec, task_index: int, spec: MLPSpec, opt: OptimizerConfig,
                 cfg: EngineConfig, batch_size: int = 32, device: int = 0, rdv: Optional[Rendezvous] = None,
                 layout: Optional[ModelLayout] = None, verbose: bool = False):
        cfg.validate(opt)
        if batch_size < 1 or batch_size > 256:
            raise ValueError("batch_size must be in [1, 256] per worker")
        if spec.num_classes > head_ops.MAX_CLASSES:
            raise ValueError(f"at most {head_ops.MAX_CLASSES} classes")
        self.cluster, self.task_index, self.spec, self.opt, self.cfg = cluster, task_index, spec, opt, cfg
        self.batch = batch_size
        self.B_pad = _round_up(batch_size, 16)
        self.device = device if cfg.backend == "cuda" else -1
        self.verbose = verbose
        self.layout = layout or build_layout(spec, cluster.num_ps, cfg.sharding, dw_tile_n_for(cfg.dtype))
        self.rdv = rdv or Rendezvous(cluster, "worker", task_index)
        self.lib = N.lib()
        self.is_chief = task_index == 0  # DS:108
        self.ps_segs: List[Segment] = []
        self.seg: Optional[Segment] = None
        self._exec = None
        self._connected = False
        self._closed = False
        self._seq_host = 0          # cpu backend: push sequence number
        self.dt = cfg.native_dtype
        self.es = gemm_ops.elem_size(self.dt)
        self.tdtype = torch.float32 if cfg.dtype == "fp32" else torch.bfloat16
        self.ld_in = gemm_ops.padded_ld(spec.in_features)
        # shards this worker pushes to, the one owning global_step first (its inbox entry is index 0)
        used = [sh.ps for sh in self.layout.shards if sh.n_items > 0]
        gs_owner = self.layout.placement["global_step"]
        self.gs_owner = gs_owner
        self.inbox_order = ([gs_owner] if gs_owner in used else []) + [k for k in used if k != gs_owner]
        self.inbox_index = {k: i for i, k in enumerate(self.inbox_order)}
        self.kernels_per_step = 0

    # ------------------------------------------------------------------------------------------
    # bring-up
    # ------------------------------------------------------------------------------------------
    def connect(self, timeout_s: Optional[float] = None) -> None:
        """Open every PS shard, publish our inbox, and (non-chief) wait for the chief's initialisation."""
        cfg = self.cfg
        if cfg.backend == "cuda":
            N.check(self.lib.dm_set_device(self.device), "set device")
            torch.cuda.set_device(self.device)
        self.ps_desc = []
        for k in range(self.cluster.num_ps):
            desc = self.rdv.get(f"ps/{k}/segment", timeout_s)
            if desc["nslots"] != cfg.nslots or desc["n_workers"] != self.cluster.num_workers:
                raise RuntimeError(f"ps {k} was started with a different engine configuration: {desc}")
            if desc["arena_elems"] != self.layout.shards[k].arena_elems or desc["n_items"] != self.layout.shards[k].n_items:
                raise RuntimeError(f"ps {k} has a different model layout (model / sharding flags differ?)")
            self.ps_desc.append(desc)
            self.ps_segs.append(Segment.open(desc, device=self.device))
        carver = Carver()
        carver.add("inbox", max(1, len(self.inbox_order)) * 8)
        carver.add("seq", 4 * (1 + MAX_SLOTS))   # [0] pushes opened so far, [1 + slot] seq of the slot's current step
        kind = "cuda" if cfg.backend == "cuda" else "shm"
        self.seg = Segment.create(kind, carver.total, device=self.device, table=carver.table(),
                                  tag=f"w{self.task_index}")
        desc = self.seg.export()
        desc["inbox_index"] = {str(k): i for k, i in self.inbox_index.items()}
        desc["worker_device"] = self.device
        # a worker that is restarted after a crash registers again under the next incarnation number; the ps
        # tasks then re-admit it (fresh push sequence, fresh inbox) — the reference's recoverable session
        self.incarnation = self.rdv.add(f"worker/{self.task_index}/incarnation", 1)
        desc["incarnation"] = self.incarnation
        self.rdv.put(f"worker/{self.task_index}/inbox", desc)
        self.heartbeat()
        self._connected = True

    def wait_ready(self, timeout_s: Optional[float] = None) -> None:
        """Block until the variables are initialised (DS:108-109 non-chief wait) and every PS shard serves us."""
        self.rdv.get("init/done", timeout_s)
        for k in range(self.cluster.num_ps):
            self.rdv.get(f"ps/{k}/serving", timeout_s)
            self.rdv.get(f"ps/{k}/attached/{self.task_index}/{self.incarnation}", timeout_s)
        self.prepare()
        # seed our view of the shared step counter (a restored session does not start at 0)
        g0 = self.read_global_step()
        if self.cfg.backend == "cuda":
            host = C.c_uint32(g0)
            N.check(self.lib.dm_memcpy_async(self.seg.addr("inbox", 4), C.addressof(host), 4, None))
            N.check(self.lib.dm_stream_sync(None))
        else:
            self.lib.dm_store_release_u32(self.seg.addr("inbox", 4), g0)

    def _read_own(self, region: str, count: int, byte_offset: int = 0) -> List[int]:
        if self.cfg.backend == "cuda":
            host = (C.c_uint32 * count)()
            N.check(self.lib.dm_memcpy_async(C.addressof(host), self.seg.addr(region, byte_offset), 4 * count, None))
            N.check(self.lib.dm_stream_sync(None))
            return list(host)
        return [self.lib.dm_load_acquire_u32(self.seg.addr(region, byte_offset + 4 * i)) for i in range(count)]

    def pushes_made(self) -> int:
        if self.cfg.backend == "cuda":
            self.drain()
            return self._read_own("seq", 1)[0]
        return self._seq_host

    def wait_applied(self, timeout_s: float = 60.0) -> None:
        """Block until every push this worker has made is applied on every shard (needed for a consistent
        checkpoint or evaluation; training itself never waits like this)."""
        if self.cfg.push_mode != "mailbox" or not self.inbox_order:
            self.drain()
            return
        seq = self.pushes_made()
        t0 = time.time()
        while True:
            inbox = self._read_own("inbox", 2 * len(self.inbox_order))
            if all(((inbox[2 * i] - seq) & 0xFFFFFFFF) < 0x80000000 for i in range(len(self.inbox_order))):
                return
            if time.time() - t0 > timeout_s:
                raise TimeoutError(f"pushes up to {seq} not acknowledged: inbox={inbox}")
            time.sleep(0.0005)

    def heartbeat(self) -> None:
        """Liveness mark for the ps-side failure detector (`ParameterServer.join(worker_timeout_s=...)`)."""
        self.rdv.put(f"session/heartbeat/{self.task_index}", time.time())

    def prepare(self) -> None:
        """Build the step graphs (GPU backend). Only needs the PS pointers, so it may run before the variables
        are initialised — in-process clusters call it before the persistent PS kernel is launched so that no
        allocation happens while that kernel owns part of the GPU."""
        if self.cfg.backend == "cuda" and self._exec is None:
            self._build_cuda()

    # ------------------------------------------------------------------------------------------
    # variable I/O through peer memory (chief init / restore, checkpoints, evaluation)
    # ------------------------------------------------------------------------------------------
    def _copy_to_ps(self, k: int, region: str, elem_offset: int, src: torch.Tensor) -> None:
        src = src.contiguous()
        nbytes = src.numel() * src.element_size()
        dst = self.ps_segs[k].addr(region, elem_offset * src.element_size())
        if self.cfg.backend == "cuda":
            host = src.cpu()
            N.check(self.lib.dm_memcpy_async(dst, host.data_ptr(), nbytes, None), "copy to ps")
            N.check(self.lib.dm_stream_sync(None))
        else:
            C.memmove(dst, src.data_ptr(), nbytes)

    def _copy_from_ps(self, k: int, region: str, elem_offset: int, count: int, dtype: torch.dtype) -> torch.Tensor:
        out = torch.empty(count, dtype=dtype)
        nbytes = count * out.element_size()
        src = self.ps_segs[k].addr(region, elem_offset * out.element_size())
        if self.cfg.backend == "cuda":
            N.check(self.lib.dm_memcpy_async(out.data_ptr(), src, nbytes, None), "copy from ps")
            N.check(self.lib.dm_stream_sync(None))
        else:
            C.memmove(out.data_ptr(), src, nbytes)
        return out

    def _pack(self, vl: VarLayout, t: torch.Tensor) -> torch.Tensor:
        """Variable tensor ([out, in] or [out]) -> arena span (row padding applied)."""
        t = t.detach().to(torch.float32).cpu()
        if t.dim() == 2 and vl.ld != vl.cols:
            buf = torch.zeros(vl.rows, vl.ld)
            buf[:, :vl.cols] = t
            return buf.reshape(-1)
        return t.reshape(-1).contiguous()

    def _unpack(self, vl: VarLayout, flat: torch.Tensor) -> torch.Tensor:
        if len(vl.spec.shape) == 2:
            return flat.view(vl.rows, vl.ld)[:, :vl.cols].clone()
        return flat.clone()

    def write_variables(self, params: Dict[str, torch.Tensor], region: str = "params") -> None:
        for name, t in params.items():
            vl = self.layout.by_name[name]
            if tuple(t.shape) != tuple(vl.spec.shape):
                raise ValueError(f"{name}: expected shape {vl.spec.shape}, got {tuple(t.shape)}")
            flat = self._pack(vl, t)
            self._copy_to_ps(vl.ps, region, vl.offset, flat)
            if region == "params" and self.cfg.dtype == "bf16":
                self._copy_to_ps(vl.ps, "shadow", vl.offset, flat.to(torch.bfloat16))

    def read_variables(self, region: str = "params") -> Dict[str, torch.Tensor]:
        out = {}
        for name, vl in self.layout.by_name.items():
            flat = self._copy_from_ps(vl.ps, region, vl.offset, vl.span, torch.float32)
            out[name] = self._unpack(vl, flat)
        return out

    def read_global_step(self) -> int:
        t = self._copy_from_ps(self.gs_owner, "ctrl", CTRL_GLOBAL_STEP, 1, torch.int32)
        return int(t.item()) & 0xFFFFFFFF

    def write_global_step(self, value: int) -> None:
        self._copy_to_ps(self.gs_owner, "ctrl", CTRL_GLOBAL_STEP, torch.tensor([value], dtype=torch.int32))

    def read_item_state(self, k: int) -> torch.Tensor:
        ni = max(self.layout.shards[k].n_items, 1)
        return self._copy_from_ps(k, "item_state", 0, ni * 4, torch.int32)

    def write_item_state(self, k: int, state: torch.Tensor) -> None:
        self._copy_to_ps(k, "item_state", 0, state.to(torch.int32))

    def initialize_variables(self, seed: int = 0, params: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Chief-only: run the initialisers on the PS shards and announce `init/done` (DS:108-109)."""
        if not self.is_chief:
            raise RuntimeError("only the chief (worker 0) initialises variables")
        params = params if params is not None else mlp.init_params(self.spec, seed)
        self.write_variables(params)
        zeros = {k: torch.zeros_like(v) for k, v in params.items()}
        self.write_variables(zeros, "adam_m")
        self.write_variables(zeros, "adam_v")
        self.rdv.put("init/done", {"seed": seed, "by": self.task_index})

    def mark_initialized(self) -> None:
        """Chief-only: announce that the variables on the PS are valid (after a checkpoint restore)."""
        self.rdv.put("init/done", {"restored": True, "by": self.task_index})

    # ------------------------------------------------------------------------------------------
    # GPU step construction
    # ------------------------------------------------------------------------------------------
    def _push_target(self, k: int, seq_ptr: int) -> N.PushTarget:
        seg, desc = self.ps_segs[k], self.ps_desc[k]
        t = N.PushTarget()
        w = self.task_index
        arena, ni, ns = desc["arena_elems"], max(desc["n_items"], 1), self.cfg.nslots
        if self.cfg.push_mode == "atomic":
            t.mode, t.scale, t.base = N.PUSH_ATOMIC, -self.opt.lr, seg.addr("params")
            t.nslots = 1
        else:
            t.mode, t.scale = N.PUSH_MAILBOX, 1.0
            t.base = seg.addr("mailbox", w * ns * arena * 4)
            t.slot_stride = arena
            t.flags = seg.addr("flags", w * ns * ni * 4)
            t.flag_slot_stride = ni
            t.nslots = ns
        t.seq_ptr = seq_ptr
        # ps shard on our own GPU (in-process / colocated): gpu-scope release suffices for its flags
        t.gpu_scope = int(self.cfg.backend == "cuda" and desc.get("device", -2) == self.device
                          and self.cluster.num_workers == 1)
        return t

    def _weight_ptr(self, vl: VarLayout) -> int:
        seg = self.ps_segs[vl.ps]
        if self.cfg.dtype == "bf16":
            return seg.addr("shadow", vl.offset * 2)
        return seg.addr("params", vl.offset * 4)

    def _build_cuda(self) -> None:
        spec, cfg, lay = self.spec, self.cfg, self.layout
        N.ensure_prepared(self.device)
        x_bytes = self.B_pad * self.ld_in * self.es
        y_bytes = self.B_pad * spec.num_classes * 4
        out = C.c_void_p()
        N.check(self.lib.dm_exec_create(self.device, cfg.pipeline_slots, cfg.lanes, cfg.graph_steps, x_bytes, y_bytes,
                                        C.byref(out)), "exec create")
        self._exec = out.value
        self.x_bytes, self.y_bytes = x_bytes, y_bytes
        dev = f"cuda:{self.device}"
        names = spec.variable_names()
        sizes = spec.layer_sizes
        L = len(sizes)
        # activations / pre-activation gradients of the hidden layers (row padded, zero initialised)
        # (one set per executor slot: steps of different slots may run concurrently)
        self._slot_act = [[None] + [torch.zeros(self.B_pad, gemm_ops.padded_ld(sizes[l][1]), dtype=self.tdtype,
                                                device=dev) for l in range(L - 1)] for _ in range(cfg.pipeline_slots)]
        self._slot_dact = [[None] + [torch.zeros_like(a[l + 1]) for l in range(L - 1)] for a in self._slot_act]
        self.act, self.dact = self._slot_act[0], self._slot_dact[0]
        seq_counter = self.seg.addr("seq")
        self._slots = []
        stream = self.lib.dm_exec_compute_stream(self._exec)
        # ---- evaluation path (forward + head without gradients; shares the activation buffers) ----
        self._eval_x = torch.zeros(self.B_pad, self.ld_in, dtype=self.tdtype, device=dev)
        self._eval_y = torch.zeros(self.B_pad, spec.num_classes, dtype=torch.float32, device=dev)
        self._eval_res = torch.zeros(4, dtype=torch.int32, device=dev)
        self._eval_plans = []
        for l in range(L - 1):
            wl, bl = lay.by_name[names[l][0]], lay.by_name[names[l][1]]
            fin, fout = sizes[l]
            self._eval_plans.append(gemm_ops.forward_plan(
                w_ptr=self._weight_ptr(wl), x_ptr=self._eval_x.data_ptr() if l == 0 else self.act[l].data_ptr(),
                out_ptr=self.act[l + 1].data_ptr(), bias_ptr=self.ps_segs[bl.ps].addr("params", bl.offset * 4),
                O=fout, I=fin, B=self.batch, B_pad=self.B_pad, dtype=self.dt, relu=True, ldw=wl.ld,
                ldx=self.ld_in if l == 0 else self.act[l].shape[1], ldo=self.act[l + 1].shape[1], name=f"eval_fwd{l}"))
        wl, bl = lay.by_name[names[L - 1][0]], lay.by_name[names[L - 1][1]]
        self._eval_plans.append(head_ops.head_plan(
            h_ptr=self.act[L - 1].data_ptr(), labels_ptr=self._eval_y.data_ptr(),
            w_last_ptr=self.ps_segs[wl.ps].addr("params", wl.offset * 4),
            b_last_ptr=self.ps_segs[bl.ps].addr("params", bl.offset * 4), dpre_ptr=self.dact[L - 1].data_ptr(),
            result_ptr=self._eval_res.data_ptr(), B=self.batch, B_pad=self.B_pad, H=sizes[L - 1][0],
            num_classes=spec.num_classes, loss_kind=N.LOSS_BOOK if spec.loss == "book" else N.LOSS_XENT,
            act_bf16=cfg.dtype == "bf16", compute_grads=False, ldh=self.act[L - 1].shape[1]))
        torch.cuda.synchronize(self.device)
        for slot in range(cfg.pipeline_slots):
            act, dact = self._slot_act[slot], self._slot_dact[slot]
            seq_ptr = self.seg.addr("seq", 4 * (1 + slot))
            stream = self.lib.dm_exec_capture_stream(self._exec, slot)
            xd, yd, rd, xs, ys = (C.c_void_p() for _ in range(5))
            N.check(self.lib.dm_exec_slot_info(self._exec, slot, C.byref(xd), C.byref(yd), C.byref(rd), C.byref(xs),
                                               C.byref(ys)))
            plans = []
            # ---- forward: hidden layers, W pulled from the PS shard by TMA ----
            for l in range(L - 1):
                wn, bn = names[l]
                wl, bl = lay.by_name[wn], lay.by_name[bn]
                fin, fout = sizes[l]
                in_ptr = xd.value if l == 0 else act[l].data_ptr()
                ld_in = self.ld_in if l == 0 else act[l].shape[1]
                plans.append(gemm_ops.forward_plan(
                    w_ptr=self._weight_ptr(wl), x_ptr=in_ptr, out_ptr=act[l + 1].data_ptr(),
                    bias_ptr=self.ps_segs[bl.ps].addr("params", bl.offset * 4), O=fout, I=fin, B=self.batch,
                    B_pad=self.B_pad, dtype=self.dt, relu=True, ldw=wl.ld, ldx=ld_in, ldo=act[l + 1].shape[1],
                    bump_seq_ptr=seq_ptr if l == 0 else 0, seq_counter_ptr=seq_counter, name=f"fwd{l}"))
            # ---- classifier head ----
            wn, bn = names[L - 1]
            wl, bl = lay.by_name[wn], lay.by_name[bn]
            hb = lay.by_name[names[L - 2][1]]  # bias of the last hidden layer
            inbox_ptr = self.seg.addr("inbox") if cfg.push_mode == "mailbox" else 0
            gs_ptr = self.ps_segs[self.gs_owner].addr("ctrl", 4 * CTRL_GLOBAL_STEP) if cfg.push_mode == "atomic" else 0
            head = head_ops.head_plan(
                h_ptr=act[L - 1].data_ptr(), labels_ptr=yd.value,
                w_last_ptr=self.ps_segs[wl.ps].addr("params", wl.offset * 4),
                b_last_ptr=self.ps_segs[bl.ps].addr("params", bl.offset * 4),
                dpre_ptr=dact[L - 1].data_ptr(), result_ptr=rd.value, B=self.batch, B_pad=self.B_pad,
                H=sizes[L - 1][0], num_classes=spec.num_classes,
                loss_kind=N.LOSS_BOOK if spec.loss == "book" else N.LOSS_XENT, act_bf16=cfg.dtype == "bf16",
                push=self._push_target(wl.ps, seq_ptr), push_bh=self._push_target(hb.ps, seq_ptr),
                push_bl=self._push_target(bl.ps, seq_ptr),
                off_w_last=wl.offset, off_b_last=bl.offset, off_b_hidden=hb.offset,
                item_w_last_base=wl.item_base, item_b_last=bl.item_base, item_b_hidden_base=hb.item_base,
                seq_ptr=seq_ptr, inbox_ptr=inbox_ptr, n_inbox=len(self.inbox_order), ps_global_step_ptr=gs_ptr,
                nslots=cfg.nslots, ldh=act[L - 1].shape[1])
            if cfg.fuse_head and L == 2 and plans[0].can_fuse_head(sizes[L - 1][0], spec.num_classes):
                plans[0].fuse_head(head.params)      # the head runs as the tail of the forward GEMM's cluster
            else:
                plans.append(head)
            # ---- backward of the hidden layers: dX (+ bias-grad push) *before* dW (fused push) of the same layer:
            #      dX pulls W_l from the PS a second time, and the PS applies a pushed dW_l within microseconds,
            #      so pushing dW_l first would let this step's own update leak into its dX ----
            for l in range(L - 2, -1, -1):
                wn, bn = names[l]
                wl = lay.by_name[wn]
                fin, fout = sizes[l]
                in_ptr = xd.value if l == 0 else act[l].data_ptr()
                ld_in = self.ld_in if l == 0 else act[l].shape[1]
                if l > 0:
                    pb = lay.by_name[names[l - 1][1]]  # bias of the previous hidden layer gets its grad here
                    plans.append(gemm_ops.dx_plan(
                        w_ptr=self._weight_ptr(wl), dy_ptr=dact[l + 1].data_ptr(), out_ptr=dact[l].data_ptr(),
                        mask_ptr=act[l].data_ptr(), O=fout, I=fin, B=self.batch, B_pad=self.B_pad, dtype=self.dt,
                        ldw=wl.ld, lddy=dact[l + 1].shape[1], ldo=dact[l].shape[1],
                        colsum=self._push_target(pb.ps, seq_ptr), colsum_offset=pb.offset, colsum_item_base=pb.item_base,
                        name=f"dx{l}"))
                plans.append(gemm_ops.dw_plan(
                    dy_ptr=dact[l + 1].data_ptr(), x_ptr=in_ptr, O=fout, I=fin, B_pad=self.B_pad, dtype=self.dt,
                    push=self._push_target(wl.ps, seq_ptr), push_offset=wl.offset, item_base=wl.item_base,
                    bn=lay.dw_tile_n, lddy=dact[l + 1].shape[1], ldx=ld_in, ldw=wl.ld, name=f"dw{l}"))
            if cfg.pdl:
                # programmatic dependent launch inside the step graph: kernel k+1's prologue (and the head's
                # W_last fetch) overlaps kernel k; every kernel waits on griddepcontrol.wait before it touches
                # its predecessor's outputs. The first kernel keeps a full dependency on the previous step.
                for p in plans[1:]:
                    p.params.pdl = 1
            self.kernels_per_step = len(plans)
            N.check(self.lib.dm_exec_begin_capture(self._exec, slot), "begin capture")
            try:
                for p in plans:
                    p.launch(stream)
            finally:
                N.check(self.lib.dm_exec_end_capture(self._exec, slot, len(plans)), "end capture")
            x_stage = torch.from_numpy(np.frombuffer((C.c_uint8 * x_bytes).from_address(xs.value), dtype=np.uint8))
            y_stage = torch.from_numpy(np.frombuffer((C.c_uint8 * y_bytes).from_address(ys.value), dtype=np.uint8))
            self._slots.append({
                "plans": plans, "x_dev": xd.value, "y_dev": yd.value, "res_dev": rd.value,
                "x_stage": x_stage.view(self.tdtype).view(self.B_pad, self.ld_in), "x_stage_ptr": xs.value,
                "y_stage": y_stage.view(torch.float32).view(self.B_pad, spec.num_classes), "y_stage_ptr": ys.value,
            })

        # group graphs: the U steps of a group as parallel chains of one graph (native loops launch these)
        U = cfg.graph_steps
        if U > 1:
            for g in range(cfg.pipeline_slots // U):
                N.check(self.lib.dm_exec_begin_group_capture(self._exec, g), "begin group capture")
                try:
                    for u in range(U):
                        slot = g * U + u
                        st = self.lib.dm_exec_capture_stream(self._exec, slot)
                        for p in self._slots[slot]["plans"]:
                            p.launch(st)
                finally:
                    N.check(self.lib.dm_exec_end_group_capture(self._exec, g), "end group capture")

    # ------------------------------------------------------------------------------------------
    # stepping
    # ------------------------------------------------------------------------------------------
    def submit(self, x: torch.Tensor, y: torch.Tensor) -> int:
        """Enqueue one training step on a host batch (x [B, in] float, y [B, classes] one-hot). Returns a ticket."""
        if self.cfg.backend != "cuda":
            raise RuntimeError("submit/result pipelining exists on the cuda backend; use step() on cpu")
        slot = C.c_int()
        N.check(self.lib.dm_exec_acquire_slot(self._exec, C.byref(slot)), "acquire slot")
        s = self._slots[slot.value]
        b = x.shape[0]
        if b != self.batch:
            raise ValueError(f"expected a batch of {self.batch}, got {b}")
        s["x_stage"][:b, :x.shape[1]].copy_(x)
        s["y_stage"][:b].copy_(y)
        t = C.c_uint64()
        N.check(self.lib.dm_exec_submit(self._exec, s["x_stage_ptr"], s["y_stage_ptr"], C.byref(t)), "submit")
        return t.value

    def submit_resident(self, x_ptr: int = 0, y_ptr: int = 0) -> int:
        """Enqueue a step whose inputs already live in device (or pinned) memory at x_ptr / y_ptr, laid out
        exactly like the slot buffers ([B_pad][ld_in] compute dtype, [B_pad][classes] fp32). 0 = reuse the
        data already in the slot."""
        t = C.c_uint64()
        N.check(self.lib.dm_exec_submit(self._exec, x_ptr or None, y_ptr or None, C.byref(t)), "submit")
        return t.value

    def run_resident(self, n_steps: int, x_base_ptr: int, y_base_ptr: int, x_row_bytes: int, y_row_bytes: int,
                     n_rows: int, start: int = 0) -> None:
        """Native loop over a device-resident dataset: step i trains on rows ((start + i) * batch) % n_rows ..
        (contiguous), copied device-to-device into the slot buffers. Results stay in the executor's history."""
        N.check(self.lib.dm_exec_run_resident(self._exec, n_steps, x_base_ptr, y_base_ptr, x_row_bytes, y_row_bytes,
                                              n_rows, self.batch, start), "run resident")

    def result(self, ticket: int, wait: bool = True) -> Optional[StepOutput]:
        r = N.StepResult()
        rc = self.lib.dm_exec_result(self._exec, ticket, C.addressof(r), int(wait))
        if rc == 1:
            return None
        if rc != 0:
            N.check(rc if rc < 0 else -rc, "result")
        return StepOutput(r.loss, r.global_step, r.correct, r.seq)

    def step(self, x: torch.Tensor, y: torch.Tensor) -> StepOutput:
        """One synchronous training step — the analogue of `sess.run([train_op, loss, global_step], feed_dict)`."""
        if self.cfg.backend == "cuda":
            return self.result(self.submit(x, y), wait=True)
        return self._cpu_step(x, y)

    def make_loader(self, images: torch.Tensor, labels: torch.Tensor, seed: int = 0, shuffle: bool = True):
        """Native `next_batch` loader over a host dataset (kept alive by the returned object)."""
        return NativeLoader(self, images, labels, seed, shuffle)

    def run_steps(self, n_steps: int, loader: "NativeLoader", stop_at_global_step: int = 0) -> Sequence[StepOutput]:
        """Native train loop: n_steps x (next_batch -> H2D -> step graph -> result D2H)."""
        if self.cfg.backend != "cuda":
            out = []
            for _ in range(n_steps):
                x, y = loader.next_batch()
                r = self._cpu_step(x, y)
                out.append(r)
                if stop_at_global_step and r.global_step >= stop_at_global_step:
                    break
            return out
        res = (N.StepResult * n_steps)()
        done = C.c_uint64()
        N.check(self.lib.dm_exec_run(self._exec, loader.handle, n_steps, C.addressof(res), stop_at_global_step,
                                     C.byref(done)), "exec run")
        raw = np.frombuffer(res, dtype=_STEP_DTYPE, count=n_steps)[: done.value]   # keeps `res` alive
        return StepOutputs(raw)

    def drain(self) -> None:
        if self._exec:
            N.check(self.lib.dm_exec_drain(self._exec), "drain")

    @property
    def compute_stream(self) -> int:
        """Raw cudaStream_t of the executor's compute stream (for CUDA-event timing of a run of steps)."""
        return self.lib.dm_exec_compute_stream(self._exec)

    def enqueue_wait_ack(self) -> None:
        """Stream-ordered fence: the compute stream does not proceed until the PS has applied every push made so
        far (mailbox mode). Used to close a device-timed region on the PS-side apply of its last step."""
        if self.cfg.backend != "cuda":
            return
        N.check(self.lib.dm_exec_join(self._exec), "join lanes")   # lane 0 now follows every in-flight step
        if self.cfg.push_mode == "mailbox" and self.inbox_order:
            N.check(self.lib.dm_launch_wait_ack(self.seg.addr("inbox"), len(self.inbox_order), self.seg.addr("seq"),
                                                self.compute_stream), "wait_ack")

    def fork_lanes(self) -> None:
        """Every lane waits for what is enqueued on lane 0 (e.g. a timing event) before running further steps."""
        if self.cfg.backend == "cuda" and self._exec:
            N.check(self.lib.dm_exec_fork(self._exec), "fork lanes")

    def kernel_launches(self) -> int:
        return int(self.lib.dm_exec_kernel_launches(self._exec)) if self._exec else 0

    # ------------------------------------------------------------------------------------------
    # CPU backend step (same protocol, torch math)
    # ------------------------------------------------------------------------------------------
    def _cpu_views(self):
        if getattr(self, "_views", None) is None:
            self._views = []
            for k, seg in enumerate(self.ps_segs):
                d = self.ps_desc[k]
                arena, ni = d["arena_elems"], max(d["n_items"], 1)
                w, ns = self.task_index, self.cfg.nslots
                self._views.append({
                    "params": seg.tensor("params", torch.float32),
                    "mailbox": seg.tensor("mailbox", torch.float32).view(self.cluster.num_workers, ns, arena)[w],
                    "flags_addr": seg.addr("flags", w * ns * ni * 4),
                    "ni": ni,
                })
            self._inbox_addr = self.seg.addr("inbox")
        return self._views

    def _cpu_step(self, x: torch.Tensor, y: torch.Tensor) -> StepOutput:
        views = self._cpu_views()
        lay = self.layout
        # pull (X3): snapshot of the live shard memory; concurrent applies may tear it (Hogwild, like the reference)
        params = {}
        for name, vl in lay.by_name.items():
            flat = views[vl.ps]["params"][vl.offset: vl.offset + vl.span]
            params[name] = self._unpack(vl, flat)
        loss, grads, logits = mlp.manual_loss_and_grads(self.spec, params, x.float(), y.float())
        correct = mlp.accuracy_count(logits, y)
        self._seq_host += 1
        seq, ns = self._seq_host, self.cfg.nslots
        slot = seq % ns
        if seq > ns:  # flow control: wait until push (seq - nslots) has been applied everywhere
            for k, i in self.inbox_index.items():
                if self.lib.dm_wait_ge_u32(self._inbox_addr + 8 * i, seq - ns, 120.0) != 0:
                    raise TimeoutError(f"ps {k} did not acknowledge push {seq - ns}")
        # push (X4): gradients into our mailbox slot, then publish every item's flag
        for name, vl in lay.by_name.items():
            views[vl.ps]["mailbox"][slot, vl.offset: vl.offset + vl.span] = self._pack(vl, grads[name])
        for k in self.inbox_order:
            v = views[k]
            for item in range(lay.shards[k].n_items):
                self.lib.dm_store_release_u32(v["flags_addr"] + 4 * (slot * v["ni"] + item), seq)
        if self.inbox_order:
            ack = self.lib.dm_load_acquire_u32(self._inbox_addr)
            gstep = self.lib.dm_load_acquire_u32(self._inbox_addr + 4) + (seq - ack)
        else:
            gstep = seq
        return StepOutput(float(loss), int(gstep), int(correct), seq)

    # ------------------------------------------------------------------------------------------
    # evaluation (forward only, accuracy kernel / torch on cpu)
    # ------------------------------------------------------------------------------------------
    def evaluate(self, images: torch.Tensor, labels: torch.Tensor) -> Tuple[float, float]:
        """(mean loss, accuracy) of the current PS variables on a host dataset; GPU path uses the hand-written
        accuracy reduction kernel on torch-computed logits of the pulled variables."""
        if self.cfg.backend == "cuda":
            # forward tcgen05 GEMMs (weights pulled from the PS shards) + the head kernel in eval mode; whole
            # batches only (the step kernels are specialised for this worker's batch size).
            self.drain()
            stream = self.compute_stream
            n_chunks = images.shape[0] // self.batch
            if n_chunks == 0:
                raise ValueError(f"evaluate needs at least one full batch of {self.batch} samples")
            xs = torch.zeros(self.B_pad, self.ld_in, dtype=self.tdtype).pin_memory()
            ys = torch.zeros(self.B_pad, self.spec.num_classes, dtype=torch.float32).pin_memory()
            res = torch.zeros(4, dtype=torch.int32).pin_memory()
            loss_sum, correct = 0.0, 0
            for c in range(n_chunks):
                sl = slice(c * self.batch, (c + 1) * self.batch)
                xs[: self.batch, : images.shape[1]].copy_(images[sl])
                ys[: self.batch].copy_(labels[sl])
                N.check(self.lib.dm_memcpy_async(self._eval_x.data_ptr(), xs.data_ptr(), xs.numel() * xs.element_size(), stream))
                N.check(self.lib.dm_memcpy_async(self._eval_y.data_ptr(), ys.data_ptr(), ys.numel() * 4, stream))
                for p in self._eval_plans:
                    p.launch(stream)
                N.check(self.lib.dm_memcpy_async(res.data_ptr(), self._eval_res.data_ptr(), 16, stream))
                N.check(self.lib.dm_stream_sync(stream))
                loss_sum += float(res[:1].view(torch.float32).item())
                correct += int(res[2].item())
            return loss_sum / n_chunks, correct / (n_chunks * self.batch)
        params = self.read_variables()
        logits, _ = mlp.forward_logits(self.spec, params, images.float())
        correct = mlp.accuracy_count(logits, labels)
        loss = float(mlp.loss_from_logits(self.spec, logits, labels.float()))
        return loss, correct / max(1, images.shape[0])

    # ------------------------------------------------------------------------------------------
    # teardown
    # ------------------------------------------------------------------------------------------
    def finish(self) -> None:
        """Leave the session: tell every PS shard how many pushes we made so its serve kernel can retire us."""
        if not self._connected or getattr(self, "_finished", False):
            return
        self._finished = True
        w = self.task_index
        if self.cfg.push_mode == "mailbox":
            if self.cfg.backend == "cuda" and self._exec:
                self.drain()
                stream = self.lib.dm_exec_compute_stream(self._exec)
                for k in self.inbox_order:
                    N.check(self.lib.dm_launch_worker_done(self.ps_segs[k].addr("ctrl", 4 * (CTRL_WORKER_DONE + w)),
                                                           self.seg.addr("seq"), stream))
                N.check(self.lib.dm_stream_sync(stream))
            elif self.cfg.backend == "cpu":
                for k in self.inbox_order:
                    self.lib.dm_store_release_u32(self.ps_segs[k].addr("ctrl", 4 * (CTRL_WORKER_DONE + w)),
                                                  self._seq_host + 1)
        self.rdv.put(f"session/done/{w}", 1)
        self.rdv.add("session/workers_done", 1)

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self.finish()
        if self._exec:
            self.lib.dm_exec_destroy(self._exec)
            self._exec = None
        self._views = None
        for seg in self.ps_segs:
            seg.close()
        if self.seg is not None:
            self.seg.close()


class NativeLoader:
    """Host dataset + native `next_batch` gatherer (TF DataSet.next_batch semantics, csrc/executor.cu)."""

    def __init__(self, worker: Worker, images: torch.Tensor, labels: torch.Tensor, seed: int = 0, shuffle: bool = True):
        self.worker = worker
        self.images = images.to(worker.tdtype).contiguous()
        self.labels = labels.to(torch.float32).contiguous()
        if worker.cfg.backend == "cuda":
            self.images = self.images.pin_memory()
            self.labels = self.labels.pin_memory()
        n, pix = self.images.shape
        es = self.images.element_size()
        self.handle = worker.lib.dm_loader_create(
            self.images.data_ptr(), self.labels.data_ptr(), n, pix * es, self.labels.shape[1] * 4,
            worker.ld_in * es, self.labels.shape[1] * 4, worker.batch, seed, int(shuffle))
        self._xbuf = torch.zeros(worker.batch, worker.ld_in, dtype=worker.tdtype)
        self._ybuf = torch.zeros(worker.batch, self.labels.shape[1], dtype=torch.float32)

    def next_batch(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self.worker.lib.dm_loader_next(self.handle, self._xbuf.data_ptr(), self._ybuf.data_ptr())
        return self._xbuf[:, : self.images.shape[1]].clone(), self._ybuf.clone()

    @property
    def epochs(self) -> int:
        return int(self.worker.lib.dm_loader_epochs(self.handle))

    def __del__(self):
        try:
            if self.handle:
                self.worker.lib.dm_loader_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
