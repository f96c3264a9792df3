"""Variable -> parameter-server placement and the per-shard memory layout.

Reference parity: `tf.device(tf.train.replica_device_setter(worker_device=..., cluster=cluster))`
(/root/reference/distributed_server-basic.py:88-89) pins every Variable — `global_step` first (DS:91), then the
model variables in creation order (DS:41-47), with the Adam slots colocated — to ps tasks round-robin, whole
variables only. `strategy="round_robin"` reproduces that order exactly (2 ps: global_step->ps0, hid_w->ps1,
hid_b->ps0, sm_w->ps1, sm_b->ps0); `strategy="byte_balanced"` is the better sharding mode for this engine
(greedy largest-first by bytes), since hid_w alone is 98.6 % of the bytes.

A shard's *arena* is one flat fp32 buffer; every variable sits at a 64-element (256-byte) aligned offset,
2-D weights of hidden layers use a padded leading dimension so TMA row strides are 16-byte multiples. The
unit of push/apply hand-off is an *item* (2-D block of the arena) — see csrc/protocol.h.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from ..models.mlp import MLPSpec, VarSpec

TILE_M = 128          # rows of a dW tile / width of a bias or head slice
DW_TILE_N = 64        # columns of a dW tile (bn of the dW GEMM)
ALIGN_ELEMS = 64      # 256-byte alignment of every variable inside the arena
GLOBAL_STEP = "global_step"


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def padded_ld(n: int) -> int:
    return _round_up(n, 8)


def place_variables(spec: MLPSpec, num_ps: int, strategy: str = "round_robin") -> Dict[str, int]:
    """Map variable name (incl. `global_step`) -> ps task index."""
    names = [GLOBAL_STEP] + [v.name for v in spec.variables()]
    if num_ps <= 0:
        raise ValueError("need at least one ps task")
    if strategy == "round_robin":
        return {n: i % num_ps for i, n in enumerate(names)}
    if strategy == "byte_balanced":
        sizes = {v.name: v.numel * 4 for v in spec.variables()}
        sizes[GLOBAL_STEP] = 4
        load = [0] * num_ps
        out: Dict[str, int] = {GLOBAL_STEP: 0}  # the step counter stays on ps 0 (workers read it from there)
        load[0] += 4
        for n in sorted((n for n in names if n != GLOBAL_STEP), key=lambda k: -sizes[k]):
            k = min(range(num_ps), key=lambda j: (load[j], j))
            out[n] = k
            load[k] += sizes[n]
        return out
    raise ValueError(f"unknown sharding strategy {strategy!r} (round_robin | byte_balanced)")


@dataclass(frozen=True)
class Item:
    offset: int   # element offset in the arena
    rows: int
    cols: int
    ld: int
    shadow: bool  # refresh the bf16 shadow copy of this block on apply


@dataclass
class VarLayout:
    spec: VarSpec
    ps: int
    offset: int          # element offset of the variable inside its shard's arena
    ld: int              # leading dimension (== cols for vectors / the classifier weight)
    item_base: int       # index of this variable's first item in the shard's item table
    n_items: int
    role: str            # "hidden_w" | "hidden_b" | "last_w" | "last_b"

    @property
    def rows(self) -> int:
        return self.spec.shape[0] if len(self.spec.shape) == 2 else 1

    @property
    def cols(self) -> int:
        return self.spec.shape[-1]

    @property
    def span(self) -> int:
        """Elements the variable occupies in the arena (with row padding)."""
        return self.rows * self.ld if len(self.spec.shape) == 2 else self.cols


@dataclass
class ShardLayout:
    ps: int
    variables: List[VarLayout] = field(default_factory=list)
    items: List[Item] = field(default_factory=list)
    arena_elems: int = 0
    owns_global_step: bool = False

    @property
    def n_items(self) -> int:
        return len(self.items)


@dataclass
class ModelLayout:
    spec: MLPSpec
    placement: Dict[str, int]
    shards: List[ShardLayout]
    by_name: Dict[str, VarLayout]
    dw_tile_n: int = DW_TILE_N     # columns of a dW tile == bn of the dW GEMMs == width of a hidden-weight item

    def shard_of(self, name: str) -> ShardLayout:
        return self.shards[self.by_name[name].ps]


def dw_tile_n_for(dtype: str) -> int:
    """dW tile width per compute dtype. fp32 (tf32 MMA, 32-element = 128-byte slabs) uses 32-column tiles: twice
    the tiles of the 64-column layout, i.e. twice the worker CTAs pushing and twice the ps CTAs applying a push
    (the per-item apply is what bounds a shard's push rate). bf16 slabs are 64 elements wide."""
    return 32 if dtype == "fp32" else 64


def _items_for(role: str, offset: int, rows: int, cols: int, ld: int, dw_tile_n: int = DW_TILE_N) -> List[Item]:
    items: List[Item] = []
    if role == "hidden_w":
        # tile order == gridDim of the dW GEMM: mtile-major, then ntile (tile = mtile * ntiles + ntile)
        for r0 in range(0, rows, TILE_M):
            for c0 in range(0, cols, dw_tile_n):
                items.append(Item(offset + r0 * ld + c0, min(TILE_M, rows - r0), min(dw_tile_n, cols - c0), ld, True))
    elif role == "hidden_b":
        for c0 in range(0, cols, TILE_M):
            items.append(Item(offset + c0, 1, min(TILE_M, cols - c0), cols, False))
    elif role == "last_w":
        # head kernel: CTA j pushes columns [128 j, 128 j + 128) of every class row
        for c0 in range(0, cols, TILE_M):
            items.append(Item(offset + c0, rows, min(TILE_M, cols - c0), ld, False))
    elif role == "last_b":
        items.append(Item(offset, 1, cols, cols, False))
    else:
        raise ValueError(role)
    return items


def build_layout(spec: MLPSpec, num_ps: int, strategy: str = "round_robin", dw_tile_n: int = DW_TILE_N) -> ModelLayout:
    placement = place_variables(spec, num_ps, strategy)
    shards = [ShardLayout(ps=k) for k in range(num_ps)]
    shards[placement[GLOBAL_STEP]].owns_global_step = True
    by_name: Dict[str, VarLayout] = {}
    n_layers = len(spec.layer_sizes)
    for v in spec.variables():
        k = placement[v.name]
        sh = shards[k]
        last = v.layer == n_layers - 1
        if v.kind == "weight":
            role = "last_w" if last else "hidden_w"
            rows, cols = v.shape
            ld = cols if last else padded_ld(cols)
            span = rows * ld
        else:
            role = "last_b" if last else "hidden_b"
            rows, cols = 1, v.shape[0]
            ld = cols
            span = cols
        offset = _round_up(sh.arena_elems, ALIGN_ELEMS)
        items = _items_for(role, offset, rows, cols, ld, dw_tile_n)
        vl = VarLayout(spec=v, ps=k, offset=offset, ld=ld, item_base=len(sh.items), n_items=len(items), role=role)
        sh.items.extend(items)
        sh.variables.append(vl)
        sh.arena_elems = offset + span
        by_name[v.name] = vl
    for sh in shards:
        sh.arena_elems = max(_round_up(sh.arena_elems, ALIGN_ELEMS), ALIGN_ELEMS)
    return ModelLayout(spec=spec, placement=placement, shards=shards, by_name=by_name, dw_tile_n=dw_tile_n)


def shard_bytes_summary(layout: ModelLayout) -> List[Tuple[int, int, int]]:
    """[(ps index, parameter bytes, number of items)] — for logs and tests."""
    return [(sh.ps, sum(v.spec.numel for v in sh.variables) * 4, sh.n_items) for sh in layout.shards]
