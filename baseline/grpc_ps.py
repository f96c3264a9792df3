"""STAND-IN for the reference's data path (NOT the reference, which needs TensorFlow 1.x and cannot run here).

What it reproduces of `/root/reference/distributed_server-basic.py`:
  * one ps process holding the variables, Adam slots and global_step (DS:88-91, 102-103), reachable over **gRPC**
    on its `host:port` (DS:80; grpcio is the transport TF's `tf.train.Server` uses);
  * per worker step (DS:112): full-model **pull** ps->worker, forward/backward on the worker with library kernels
    (torch/cuBLAS standing in for TF's cuBLAS/Eigen ops), full-gradient **push** worker->ps, host-staged;
  * the optimizer apply runs **on the ps**, unlocked, once per push; global_step += 1 per push and is returned.

    python -m baseline.grpc_ps --workers 2 --steps 300            # spawns 1 ps + N workers on 127.0.0.1

Prints one JSON line: steps/sec over all workers (wall clock of the slowest worker, after warm-up).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from concurrent import futures

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dist_mnist_b200.models import mlp  # noqa: E402
from dist_mnist_b200.utils import data  # noqa: E402

_ident = lambda b: b  # noqa: E731  (raw-bytes (de)serialiser)


class PsState:
    def __init__(self, spec, lr, seed=0):
        self.spec = spec
        self.params = {k: v.numpy().copy() for k, v in mlp.init_params(spec, seed).items()}
        self.names = list(self.params)
        self.m = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.lr, self.b1, self.b2, self.eps = lr, 0.9, 0.999, 1e-8
        self.t = 0
        self.global_step = 0
        self.sizes = [self.params[k].size for k in self.names]

    def pull(self, _req, _ctx):
        return b"".join(self.params[k].tobytes() for k in self.names)

    def push(self, req, _ctx):
        flat = np.frombuffer(req, dtype=np.float32)
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        off = 0
        for k, n in zip(self.names, self.sizes):
            g = flat[off:off + n].reshape(self.params[k].shape)
            off += n
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            self.params[k] -= lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)
        self.global_step += 1
        return np.int64(self.global_step).tobytes()


def run_ps(args):
    import grpc

    spec = mlp.book_model(args.hidden_units)
    st = PsState(spec, args.learning_rate)
    handlers = {
        "Pull": grpc.unary_unary_rpc_method_handler(st.pull, request_deserializer=_ident, response_serializer=_ident),
        "Push": grpc.unary_unary_rpc_method_handler(st.push, request_deserializer=_ident, response_serializer=_ident),
    }
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=16),
                         options=[("grpc.max_receive_message_length", 1 << 30), ("grpc.max_send_message_length", 1 << 30)])
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("ps.Ps", handlers),))
    server.add_insecure_port(f"127.0.0.1:{args.port}")
    server.start()
    print("ps up", flush=True)
    server.wait_for_termination()   # server.join(): never returns (DS:83)


def run_worker(args):
    import grpc

    spec = mlp.book_model(args.hidden_units)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    if dev == "cuda":
        torch.cuda.set_device((1 + args.task_index) % torch.cuda.device_count())
    ch = grpc.insecure_channel(f"127.0.0.1:{args.port}",
                               options=[("grpc.max_receive_message_length", 1 << 30), ("grpc.max_send_message_length", 1 << 30)])
    grpc.channel_ready_future(ch).result(timeout=60)
    pull = ch.unary_unary("/ps.Ps/Pull", request_serializer=_ident, response_deserializer=_ident)
    push = ch.unary_unary("/ps.Ps/Push", request_serializer=_ident, response_deserializer=_ident)
    shapes = [(v.name, v.shape) for v in spec.variables()]
    ds = data.synthetic_mnist(4096, seed=args.task_index)
    it = data.BatchIterator(ds, seed=args.task_index)

    def step():
        flat = np.frombuffer(pull(b""), dtype=np.float32)
        params, off = {}, 0
        for name, shp in shapes:
            n = int(np.prod(shp))
            params[name] = torch.from_numpy(flat[off:off + n].reshape(shp).copy()).to(dev)   # host-staged pull
            off += n
        x, y = it.next_batch(args.batch_size)
        loss, grads, _ = mlp.loss_and_grads(spec, params, x.to(dev), y.to(dev))
        g = torch.cat([grads[name].reshape(-1) for name, _ in shapes]).cpu().numpy().tobytes()   # host-staged push
        gs = int(np.frombuffer(push(g), dtype=np.int64)[0])
        return float(loss), gs

    for _ in range(args.warmup):
        step()
    if dev == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss, gs = step()
    if dev == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"worker": args.task_index, "seconds": dt, "steps": args.steps, "loss": loss, "global_step": gs,
                      "device": dev}), flush=True)


def run_launcher(args):
    port = args.port
    env = dict(os.environ, PYTHONPATH=ROOT)
    base = [sys.executable, "-m", "baseline.grpc_ps", "--port", str(port), "--steps", str(args.steps), "--warmup",
            str(args.warmup), "--hidden_units", str(args.hidden_units), "--batch_size", str(args.batch_size)]
    ps = subprocess.Popen(base + ["--role", "ps"], env=env, cwd=ROOT, stdout=subprocess.PIPE, text=True)
    ps.stdout.readline()
    ws = [subprocess.Popen(base + ["--role", "worker", "--task_index", str(i)], env=env, cwd=ROOT,
                           stdout=subprocess.PIPE, text=True) for i in range(args.workers)]
    outs = []
    for w in ws:
        out, _ = w.communicate(timeout=1800)
        outs.append(json.loads(out.strip().splitlines()[-1]))
    ps.kill()
    slowest = max(o["seconds"] for o in outs)
    print(json.dumps({
        "impl": "grpc-stand-in (NOT the reference; host-staged gRPC PS with torch kernels)",
        "metric": "MNIST-MLP steps/sec (whole job, wall clock of the slowest worker)",
        "value": args.workers * args.steps / slowest, "unit": "steps/s", "workers": args.workers,
        "steps_per_worker": args.steps, "ms_per_step": slowest / args.steps * 1e3, "device": outs[0]["device"],
        "final_loss": outs[0]["loss"], "global_step": max(o["global_step"] for o in outs),
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--role", choices=["launcher", "ps", "worker"], default="launcher")
    ap.add_argument("--task_index", type=int, default=0)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--port", type=int, default=9931)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--hidden_units", type=int, default=100)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--learning_rate", type=float, default=1e-4)
    args = ap.parse_args()
    {"launcher": run_launcher, "ps": run_ps, "worker": run_worker}[args.role](args)


if __name__ == "__main__":
    main()
