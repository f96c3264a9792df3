"""Engine configuration shared by ps and worker tasks."""
from __future__ import annotations

import os
from dataclasses import dataclass, asdict

from .. import _native as N


MAX_SLOTS = 32   # executor ring depth limit (one push-sequence word per slot in the worker segment)


def _env_flag(name: str, default: bool) -> bool:
    v = os.environ.get(name)
    return default if v is None else v not in ("0", "", "false", "False")


@dataclass(frozen=True)
class OptimizerConfig:
    """Reference: `tf.train.AdamOptimizer(FLAGS.learning_rate)` with TF defaults (DS:102); SGD is the
    north-star "async-SGD" variant."""
    kind: str = "adam"          # "adam" | "sgd"
    lr: float = 1e-4            # --learning_rate default (DS:15)
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8

    @property
    def native_kind(self) -> int:
        if self.kind == "adam":
            return N.OPT_ADAM
        if self.kind == "sgd":
            return N.OPT_SGD
        raise ValueError(f"unknown optimizer {self.kind!r} (adam | sgd)")


@dataclass(frozen=True)
class EngineConfig:
    backend: str = "cuda"            # "cuda" (sm_100a kernels over NVLink peer memory) | "cpu" (shm plumbing backend)
    dtype: str = "fp32"              # compute dtype of the dense layers: "fp32" (tf32 tensor cores) | "bf16"
    nslots: int = 2                  # mailbox slots per worker (pushes in flight before the worker waits for an ack)
    apply_mode: str = "per_push"     # "per_push" (reference semantics) | "merged"
    push_mode: str = "mailbox"       # "mailbox" (PS kernel applies; Adam or SGD) | "atomic" (SGD red.add, no PS kernel)
    sharding: str = "round_robin"    # "round_robin" (reference parity) | "byte_balanced"
    ps_ctas: int = 0                 # CTAs of the persistent PS kernel (0 = auto: 120 on a dedicated ps GPU,
                                     # 32 when a worker shares the GPU)
    pipeline_slots: int = 4          # worker executor ring depth
    lanes: int = 1                   # steps of one worker in flight on the GPU at once. 1 = each step starts after
                                     # the previous one's kernels (reference-like); n = step i+1's pull/forward
                                     # overlaps step i's backward/push (asynchronous SGD; needs nslots >= lanes)
    graph_steps: int = 1             # steps per CUDA-graph launch in the native loops (U parallel step chains in
                                     # one graph: one input transfer + one launch per U steps); divides lanes
    pdl: bool = _env_flag("DM_PDL", True)   # programmatic dependent launch between the kernels of a step graph
    fuse_head: bool = _env_flag("DM_FUSED_HEAD", False)   # single-hidden-layer models: run the classifier head as
                                     # the tail of the forward GEMM's cluster (2 kernels per step instead of 3)
    colocate: bool = False           # worker i shares GPU i with ps i (N workers on N GPUs)

    @property
    def native_dtype(self) -> int:
        if self.dtype == "fp32":
            return N.DT_F32
        if self.dtype == "bf16":
            return N.DT_BF16
        raise ValueError(f"unknown dtype {self.dtype!r} (fp32 | bf16)")

    @property
    def native_apply_mode(self) -> int:
        if self.apply_mode == "per_push":
            return N.APPLY_PER_PUSH
        if self.apply_mode == "merged":
            return N.APPLY_MERGED
        raise ValueError(f"unknown apply_mode {self.apply_mode!r}")

    def validate(self, opt: OptimizerConfig) -> None:
        if self.backend not in ("cuda", "cpu"):
            raise ValueError(f"unknown backend {self.backend!r}")
        if self.push_mode not in ("mailbox", "atomic"):
            raise ValueError(f"unknown push_mode {self.push_mode!r}")
        if self.push_mode == "atomic":
            if opt.kind != "sgd":
                raise ValueError("push_mode='atomic' fuses the push with an SGD apply (red.add); use --optimizer sgd")
            if self.dtype != "fp32":
                raise ValueError("push_mode='atomic' updates only the fp32 master copy; use dtype fp32")
            if self.backend != "cuda":
                raise ValueError("push_mode='atomic' needs the cuda backend")
        if self.nslots < 1:
            raise ValueError("nslots must be >= 1")
        u = self.graph_steps
        if self.lanes < 1 or u < 1 or self.lanes % u != 0:
            raise ValueError("lanes must be a positive multiple of graph_steps")
        if self.push_mode == "mailbox" and self.lanes > self.nslots:
            raise ValueError("lanes (steps in flight) cannot exceed nslots (mailbox slots per worker)")
        if not (self.lanes <= self.pipeline_slots <= MAX_SLOTS) or self.pipeline_slots % u != 0 \
                or (self.pipeline_slots // u) % (self.lanes // u) != 0:
            raise ValueError(f"pipeline_slots must be <= {MAX_SLOTS}, a multiple of graph_steps, and hold a whole "
                             "number of rounds of lanes")
        _ = self.native_dtype, self.native_apply_mode, opt.native_kind

    def as_dict(self) -> dict:
        return asdict(self)
