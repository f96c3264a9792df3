"""Command-line entry point: the reference's flag surface on the new engine.

Reference parity (`/root/reference/distributed_server-basic.py`):
  * DS:9-24    nine flags: data_dir, hidden_units, train_steps, batch_size, learning_rate, ps_hosts,
               worker_hosts, job_name, task_index — same names and types here.
  * DS:59-67   validation + echo: `job name : {}` / `task index : {}`, `ValueError('Must specify the job name
               explicitly')`, `ValueError('Must specify a valid task index')`.
  * DS:82-83   `ps`  -> `server.join()` (never returns).
  * DS:85-116  `worker` -> build replica, chief init, train until global step 4000, print every 100 steps.

Documented divergences: the three flags that are dead in the reference are live here but default to the
values the reference *effectively* uses (`--train_steps` 4000 because of the hard-coded StopAtStepHook DS:101,
`--batch_size` 32 because of `next_batch(32)` DS:111, `--data_dir` falls back to synthetic data because the
image has no network). New flags select what the reference cannot express (optimizer, model, dtype, ...).

    python -m dist_mnist_b200.cli --job_name ps --task_index 0 --ps_hosts 127.0.0.1:9910 \
        --worker_hosts 127.0.0.1:9900,127.0.0.1:9901
"""
from __future__ import annotations

import argparse
import sys
from typing import List, Optional

import torch

from .cluster import ClusterSpec, Rendezvous, default_device_index
from .models import mlp
from .parallel.config import EngineConfig, OptimizerConfig
from .parallel.ps import ParameterServer
from .parallel.worker import Worker
from .session import train_loop
from .utils import ckpt as ckpt_utils
from .utils import data as data_utils
from .utils.metrics import TrainMetricsWriter


DEFAULT_LANES = 12   # the engine's default (bench.py uses the same); --lanes 1 reproduces the reference's sequential loop


def engine_config_from_args(args, backend: str) -> EngineConfig:
    """The engine configuration every task derives from the command line (ps and worker tasks must agree)."""
    lanes = max(1, args.lanes)
    nslots = args.nslots or max(4 * lanes, 8)
    gsteps = args.graph_steps or next(u for u in (4, 3, 2, 1) if lanes % u == 0)   # largest divisor of lanes <= 4
    return EngineConfig(backend=backend, dtype=args.dtype, nslots=nslots, apply_mode=args.apply_mode,
                        push_mode=args.push_mode, sharding=args.sharding, colocate=args.colocate, lanes=lanes,
                        graph_steps=gsteps, pipeline_slots=max(4, 2 * lanes), engine=args.engine,
                        strict_steps=args.strict_steps, ps_row_blocks=args.ps_row_blocks)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="distributed_server-basic", description=__doc__,
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    # ---- the reference's flags (DS:11-24) ----
    p.add_argument("--data_dir", type=str, default=None, help="Directory for mnist data (idx files); synthetic if absent")
    p.add_argument("--hidden_units", type=int, default=100)
    p.add_argument("--train_steps", type=int, default=4000,
                   help="stop at this *global* step (reference: flag default 10000 is dead, hook hard-codes 4000)")
    p.add_argument("--batch_size", type=int, default=32,
                   help="per-worker batch (reference: flag default 100 is dead, next_batch(32) is hard-coded)")
    p.add_argument("--learning_rate", type=float, default=0.0001)
    p.add_argument("--ps_hosts", type=str, default=None)
    p.add_argument("--worker_hosts", type=str, default=None)
    p.add_argument("--job_name", type=str, default=None, help="worker or ps")
    p.add_argument("--task_index", type=int, default=None)
    # ---- new-engine flags ----
    p.add_argument("--optimizer", choices=["adam", "sgd"], default="adam", help="reference uses Adam (DS:102)")
    p.add_argument("--adam_math", choices=["fast", "ieee"], default="fast",
                   help="ps-side Adam arithmetic (cuda): MUFU sqrt/reciprocal (default) or correctly rounded like TF's ApplyAdam")
    p.add_argument("--model", choices=["book", "zhihu", "wide"], default="book")
    p.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    p.add_argument("--backend", choices=["auto", "cuda", "cpu"], default="auto")
    p.add_argument("--push_mode", choices=["mailbox", "atomic"], default="mailbox")
    p.add_argument("--apply_mode", choices=["per_push", "merged"], default="per_push")
    p.add_argument("--sharding", choices=["round_robin", "byte_balanced", "row_split"], default="round_robin",
                   help="variable -> ps placement: round_robin = the reference's replica_device_setter; row_split also "
                        "splits the hidden weight along its input features over every ps task (fused engine)")
    p.add_argument("--engine", choices=["auto", "fused", "graph"], default="auto",
                   help="fused: one persistent kernel runs whole steps (784-H-10, H <= 128, batch <= 32, fp32); "
                        "graph: per-layer kernels in a CUDA graph (any model / dtype)")
    p.add_argument("--nslots", type=int, default=0, help="mailbox slots per worker (0 = max(4 x lanes, 8))")
    p.add_argument("--lanes", type=int, default=DEFAULT_LANES,
                   help="steps of this worker in flight at once: bounded-staleness asynchronous SGD inside one worker "
                        "(1 = strictly one step after the other, like the reference's sess.run loop)")
    p.add_argument("--strict_steps", action="store_true",
                   help="fused engine: pull the weights of a lane's next step only after the ps acknowledged its "
                        "previous push (with --lanes 1: exactly the reference's read-your-writes order)")
    p.add_argument("--graph_steps", type=int, default=0,
                   help="graph engine: steps per CUDA-graph launch in the native train loop (0 = min(lanes, 4))")
    p.add_argument("--ps_row_blocks", type=int, default=4,
                   help="fused tiling: ps items (serve-kernel CTAs) per pushed column slice of the hidden weight")
    p.add_argument("--checkpoint_dir", type=str, default=None, help="chief checkpoints here (default: mkdtemp, DS:106)")
    p.add_argument("--save_checkpoint_secs", type=float, default=600.0)
    p.add_argument("--gpu", type=int, default=None, help="CUDA device for this task (default: ps k -> k, worker i -> num_ps+i)")
    p.add_argument("--colocate", action="store_true", help="worker i shares GPU i with ps i")
    p.add_argument("--log_every", type=int, default=100)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--metrics_file", type=str, default=None,
                   help="worker: append one JSON line per --log_every local steps (loss, batch accuracy, steps/sec)")
    p.add_argument("--log_steps_per_sec", action="store_true",
                   help="worker: also print the reference's implicit StepCounterHook line (global_step/sec)")
    p.add_argument("--worker_timeout", type=float, default=0.0,
                   help="ps: declare a worker dead after this many seconds without a heartbeat and stop waiting for "
                        "it (0 = never, like the reference)")
    p.add_argument("--inject_fault", type=int, default=0,
                   help="worker: crash (os._exit) after this many local steps — fault-injection hook for tests")
    p.add_argument("--chunk_sleep", type=float, default=0.0,
                   help="worker: sleep this many seconds after every chunk of 50 steps — slows a worker down (tests of "
                        "the failure / restart path need a run that outlives a crash and a restart)")
    p.add_argument("--ps_exit_when_done", action="store_true",
                   help="let a ps task return once every worker has finished (reference ps blocks forever)")
    p.add_argument("--rendezvous_timeout", type=float, default=300.0)
    return p


def validate_task(job_name: Optional[str], task_index: Optional[int]) -> None:
    """Exact reference behaviour of DS:59-67."""
    if job_name is not None and len(job_name) > 0:
        print("job name : {}".format(job_name))
    else:
        raise ValueError("Must specify the job name explicitly")
    if task_index is not None and task_index >= 0:
        print("task index : {}".format(task_index))
    else:
        raise ValueError("Must specify a valid task index")


def resolve_backend(name: str) -> str:
    if name == "auto":
        return "cuda" if torch.cuda.is_available() else "cpu"
    return name


def run(args: argparse.Namespace) -> int:
    validate_task(args.job_name, args.task_index)
    cluster = ClusterSpec.from_flags(args.ps_hosts, args.worker_hosts)
    if args.job_name not in ("ps", "worker"):
        # the reference silently does nothing for other job names (neither branch DS:82/85 fires) after starting
        # a server that fails for an unknown job; be explicit instead.
        raise ValueError(f"unknown job name {args.job_name!r} (expected 'ps' or 'worker')")
    cluster.task_endpoint(args.job_name, args.task_index)
    backend = resolve_backend(args.backend)
    spec = mlp.get_model(args.model, args.hidden_units)
    opt = OptimizerConfig(args.optimizer, args.learning_rate, math=args.adam_math)
    cfg = engine_config_from_args(args, backend)
    cfg.validate(opt)
    device = -1
    if backend == "cuda":
        n_dev = torch.cuda.device_count()
        device = args.gpu if args.gpu is not None else default_device_index(
            cluster, args.job_name, args.task_index, n_dev, args.colocate)
    rdv = Rendezvous(cluster, args.job_name, args.task_index, timeout_s=args.rendezvous_timeout)

    if args.job_name == "ps":
        ps = ParameterServer(cluster, args.task_index, spec, opt, cfg, device=device, rdv=rdv,
                             batch_size=args.batch_size)
        ps.start()
        try:
            # DS:83 — blocks forever unless asked otherwise
            ps.join(exit_when_done=args.ps_exit_when_done, worker_timeout_s=args.worker_timeout)
        except KeyboardInterrupt:
            pass
        finally:
            try:
                print(f"[ps {args.task_index}] exiting: global_step={ps.global_step()} "
                      f"owns_global_step={int(ps.shard.owns_global_step)} items={ps.shard.n_items}", flush=True)
            except Exception:
                pass
            ps.close()
        return 0

    worker = Worker(cluster, args.task_index, spec, opt, cfg, batch_size=args.batch_size, device=device, rdv=rdv)
    dataset = data_utils.get_dataset(args.data_dir, seed=args.seed)  # DS:69 (every worker holds the full set)
    worker.connect()
    if worker.is_chief:
        if worker.variables_are_live():
            # a restarted chief (incarnation > 1): the ps tasks still hold the live training state the other workers
            # keep updating — re-running the initialisers (or restoring an older checkpoint) would reset parameters
            # and Adam slots under them while global_step and the per-item step counts keep running
            print(f"[worker 0] restarted (incarnation {worker.incarnation}): variables are live on the ps, "
                  f"not re-initialising", flush=True)
        else:
            restored = ckpt_utils.restore_latest(worker, args.checkpoint_dir)
            if restored is None:
                worker.initialize_variables(seed=args.seed)
            else:
                worker.mark_initialized()
    worker.wait_ready()
    metrics = None
    if args.metrics_file or args.log_steps_per_sec:
        metrics = TrainMetricsWriter(args.metrics_file, args.task_index, args.batch_size,
                                     every_steps=max(1, args.log_every), echo=args.log_steps_per_sec)
    try:
        res = train_loop(worker, dataset, train_steps=args.train_steps, log_every=args.log_every,
                         checkpoint_dir=args.checkpoint_dir, save_checkpoint_secs=args.save_checkpoint_secs,
                         seed=args.seed, inject_fault_after=args.inject_fault, metrics=metrics,
                         chunk_sleep_s=args.chunk_sleep)
        print(f"[worker {args.task_index}] done: local_steps={res.steps_run} last_global_step={res.last_global_step} "
              f"engine={worker.engine}", flush=True)
    finally:
        if metrics is not None:
            metrics.close()
        worker.close()
    return 0


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    return run(args)


if __name__ == "__main__":
    sys.exit(main())
