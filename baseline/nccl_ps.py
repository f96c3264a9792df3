"""NCCL baseline of the same parameter-server step (the "only calls NCCL" path the fused kernels must beat).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29621 \
        -m baseline.nccl_ps --steps 500 --warmup 20

Rank 0 is the ps (holds fp32 params + Adam state, applies with torch ops); ranks 1..N-1 are workers.
One step (synchronous, since NCCL collectives are): `broadcast(params)` ps->workers, forward/backward with
torch/cuBLAS kernels on each worker, `reduce(grads)` workers->ps (sum), ps applies Adam once per worker push
contribution (global_step += workers). With N = 1 the single rank is ps and worker (no collectives).
Device-timed with CUDA events, max over ranks; prints one JSON line from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist

from dist_mnist_b200.models import mlp
from dist_mnist_b200.utils import data


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--hidden_units", type=int, default=100)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--learning_rate", type=float, default=1e-4)
    ap.add_argument("--model", default="book")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = mlp.get_model(args.model, args.hidden_units)
    names = [v.name for v in spec.variables()]
    shapes = {v.name: v.shape for v in spec.variables()}
    sizes = [int(torch.tensor(shapes[n]).prod()) for n in names]
    init = mlp.init_params(spec, 0, device="cuda")
    flat = torch.cat([init[n].reshape(-1) for n in names]).contiguous()
    gflat = torch.zeros_like(flat)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    is_ps = rank == 0
    is_worker = world == 1 or rank > 0
    n_workers = max(1, world - 1)
    ds = data.synthetic_mnist(8192, seed=rank)
    x_all, y_all = ds.images.cuda(), ds.labels.cuda()
    B = args.batch_size
    t_adam = 0

    def views(f):
        out, off = {}, 0
        for n, s in zip(names, sizes):
            out[n] = f[off:off + s].view(shapes[n])
            off += s
        return out

    def step(i):
        nonlocal t_adam
        if world > 1:
            dist.broadcast(flat, src=0)                       # pull
        if is_worker:
            r = (i * B) % (x_all.shape[0] - B)
            params = views(flat)
            _, grads, _ = mlp.manual_loss_and_grads(spec, params, x_all[r:r + B], y_all[r:r + B])
            torch.cat([grads[n].reshape(-1) for n in names], out=gflat)
        else:
            gflat.zero_()
        if world > 1:
            dist.reduce(gflat, dst=0)                         # push (many-to-one sum)
        if is_ps:
            t_adam += 1
            lr_t = args.learning_rate * (1 - 0.999 ** t_adam) ** 0.5 / (1 - 0.9 ** t_adam)
            m.mul_(0.9).add_(gflat, alpha=0.1)
            v.mul_(0.999).addcmul_(gflat, gflat, value=0.001)
            flat.addcdiv_(m, v.sqrt().add_(1e-8), value=-lr_t)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    t1.record()
    t1.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "impl": "nccl-baseline (broadcast + reduce + torch kernels; synchronous)",
            "metric": "MNIST-MLP steps/sec (whole box, device-timed, max over ranks)",
            "value": n_workers * args.steps / (float(ms) / 1e3), "unit": "steps/s", "n_gpus": world,
            "workers": n_workers, "ms_per_step": float(ms) / args.steps,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
