"""PS push/pull bandwidth sweep 1 KB - 1 GB over NVLink peer memory vs NCCL (BASELINE.json config 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        -m bench_tools.p2p_sweep [--max_bytes 1073741824] [--out gpurun_out/p2p_sweep.json]

Rank 0 plays the ps (owns the shard buffer), ranks 1..N-1 are workers that
  push : write their buffer into their own region of the ps buffer   (in-kernel st.global / TMA bulk store)
  pull : read the ps buffer into their local buffer                   (in-kernel ld.global / TMA bulk load)
all workers concurrently (many-to-one / one-to-many through NVSwitch). Timing: CUDA events on the launching
stream, after warm-up, max over ranks; bandwidth = bytes moved per worker / time, also aggregated at the ps.
NCCL comparison: `dist.reduce` (push direction) and `dist.broadcast` (pull direction) of the same byte count.
Also measures the flag ping-pong latency between the ps GPU and worker 1.
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist

from dist_mnist_b200 import _native as N
from dist_mnist_b200.parallel.peer_mem import Carver, Segment


def time_op(fn, iters: int, warmup: int = 3) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        fn()
    t1.record()
    t1.synchronize()
    return t0.elapsed_time(t1) / iters * 1e-3


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--max_bytes", type=int, default=1 << 30)
    ap.add_argument("--min_bytes", type=int, default=1 << 10)
    ap.add_argument("--ctas", type=int, default=64)
    ap.add_argument("--out", default="gpurun_out/p2p_sweep.json")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = N.lib()
    N.ensure_prepared(local)
    n_workers = world - 1
    # ---- exchange the ps buffer through the same IPC machinery the engine uses ----
    descs = [None]
    seg = None
    if rank == 0:
        c = Carver()
        c.add("buf", args.max_bytes * max(1, n_workers))
        c.add("flags", 4096)
        seg = Segment.create("cuda", c.total, device=local, table=c.table(), tag="sweep")
        descs = [seg.export()]
    dist.broadcast_object_list(descs, src=0)
    # every rank also exports a small flag page (for the ping-pong)
    cf = Carver()
    cf.add("flag", 4096)
    myflag = Segment.create("cuda", cf.total, device=local, table=cf.table(), tag=f"flag{rank}")
    flag_descs = [None] * world
    dist.all_gather_object(flag_descs, myflag.export())
    ps = seg if rank == 0 else Segment.open(descs[0], device=local)
    local_buf = torch.empty(args.max_bytes, dtype=torch.uint8, device="cuda")
    local_buf.random_(0, 255)
    nccl_buf = torch.empty(args.max_bytes // 4, dtype=torch.float32, device="cuda")
    host_buf = torch.empty(args.max_bytes, dtype=torch.uint8).pin_memory() if rank > 0 else None
    stream = N.current_stream_ptr()
    results = []
    size = args.min_bytes
    while size <= args.max_bytes:
        iters = 50 if size <= (1 << 22) else (10 if size <= (1 << 26) else 4)
        row = {"bytes": size}
        for mode, mname in ((0, "ldst"), (1, "tma")):
            for direction in ("push", "pull"):
                t = 0.0
                if rank > 0:
                    remote = ps.addr("buf", (rank - 1) * args.max_bytes)
                    ctas = max(1, min(args.ctas, size // 16384 if mode == 1 else size // 8192)) or 1
                    if direction == "push":
                        fn = lambda: N.check(lib.dm_launch_p2p_copy(remote, local_buf.data_ptr(), size, mode, ctas, None, 0, stream))
                    else:
                        fn = lambda: N.check(lib.dm_launch_p2p_copy(local_buf.data_ptr(), remote, size, mode, ctas, None, 0, stream))
                    dist.barrier()
                    t = time_op(fn, iters)
                else:
                    dist.barrier()
                tt = torch.tensor([t], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                row[f"{direction}_{mname}_s"] = float(tt)
                row[f"{direction}_{mname}_GBps_per_worker"] = size / float(tt) / 1e9
                row[f"{direction}_{mname}_GBps_at_ps"] = n_workers * size / float(tt) / 1e9
        # NCCL equivalents: reduce to the ps (push + many-to-one sum), broadcast from the ps (pull)
        view = nccl_buf[: max(1, size // 4)]
        dist.barrier()
        t = time_op(lambda: dist.reduce(view, dst=0), iters)
        tt = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        row["nccl_reduce_s"] = float(tt)
        row["nccl_reduce_GBps_per_worker"] = size / float(tt) / 1e9
        dist.barrier()
        t = time_op(lambda: dist.broadcast(view, src=0), iters)
        tt = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        row["nccl_broadcast_s"] = float(tt)
        row["nccl_broadcast_GBps_per_worker"] = size / float(tt) / 1e9
        # Host-staged STAND-IN for the reference's gRPC data path (tf.train.Server, DS:80): the bytes leave the worker
        # GPU into pinned host memory and enter a GPU again — the two PCIe hops every gRPC tensor transfer makes —
        # WITHOUT any TCP, protobuf serialisation or host copy: an optimistic bound, labelled as a stand-in.
        t = 0.0
        if rank > 0:
            hv, dv = host_buf[:size], local_buf[:size]
            def staged():
                hv.copy_(dv, non_blocking=True)
                dv.copy_(hv, non_blocking=True)
            dist.barrier()
            t = time_op(staged, max(2, iters // 2))
        else:
            dist.barrier()
        tt = torch.tensor([t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        row["host_staged_standin_s"] = float(tt)
        row["host_staged_standin_GBps_per_worker"] = size / float(tt) / 1e9
        results.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        size *= 4
    # ---- flag ping-pong latency ps <-> worker 1 ----
    pp = None
    if world >= 2 and rank in (0, 1):
        peer = 1 - rank
        peer_flag = Segment.open(flag_descs[peer], device=local)
        out_ns = torch.zeros(1, dtype=torch.int64, device="cuda")
        iters = 2000
        torch.cuda.synchronize()
        dist.barrier(group=None) if world == 2 else None
        N.check(lib.dm_launch_pingpong(myflag.addr("flag"), peer_flag.addr("flag"), iters, rank, out_ns.data_ptr(), stream))
        torch.cuda.synchronize()
        pp = float(out_ns.item()) / iters / 2 * 1e-3  # one-way, microseconds
    if world > 2:
        dist.barrier()
    if rank == 0:
        summary = {"world": world, "workers": n_workers, "one_way_flag_latency_us": pp, "rows": results}
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(summary, f, indent=1)
        print(json.dumps({"one_way_flag_latency_us": pp}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
