"""Session helpers: bring a cluster up inside one process, and the `MonitoredTrainingSession`-style train loop.

Reference parity (`/root/reference/distributed_server-basic.py`):
  * DS:101      `StopAtStepHook(last_step=4000)` — stop once the *shared* global step reaches the limit.
  * DS:106-109  `MonitoredTrainingSession(master, is_chief, checkpoint_dir=tempfile.mkdtemp(), hooks)`:
                chief initialises / restores and checkpoints (default saver: every 600 s and at the end).
  * DS:110-116  `while not sess.should_stop(): ... if step % 100 == 0: print("Train step {}, loss: {}")`.
"""
from __future__ import annotations

import socket
import os
import tempfile
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch

from .cluster import ClusterSpec
from .utils.metrics import TrainMetricsWriter, nvtx_range
from .models.mlp import MLPSpec
from .parallel.config import EngineConfig, OptimizerConfig
from .parallel.ps import ParameterServer
from .parallel.worker import StepOutput, Worker
from .utils import ckpt as ckpt_utils
from .utils.data import Dataset


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class InProcessCluster:
    """All ps tasks plus one worker inside this process (single-GPU runs, CPU runs, tests, `smoke()`).

    On a GPU every task shares `device`; the PS serve kernel runs on its own stream next to the worker's
    step graphs, and "peer" pointers are plain local pointers.
    """

    def __init__(self, spec: MLPSpec, opt: OptimizerConfig, cfg: EngineConfig, batch_size: int = 32,
                 num_ps: int = 1, device: int = 0, seed: int = 0,
                 params: Optional[Dict[str, torch.Tensor]] = None, restore_dir: Optional[str] = None,
                 setup_hook: Optional[Callable[[Worker], None]] = None):
        base = free_port()
        ports = [base] + [free_port() for _ in range(num_ps)]
        self.cluster = ClusterSpec(tuple(f"127.0.0.1:{p}" for p in ports[:num_ps]), (f"127.0.0.1:{ports[-1]}",))
        self.ps: List[ParameterServer] = [ParameterServer(self.cluster, k, spec, opt, cfg, device=device,
                                                          batch_size=batch_size) for k in range(num_ps)]
        self.worker = Worker(self.cluster, 0, spec, opt, cfg, batch_size=batch_size, device=device)
        self.worker.connect()
        self.worker.prepare()
        if setup_hook is not None:
            setup_hook(self.worker)  # e.g. pin datasets: last chance to allocate before the PS kernel is resident
        restored = False
        if restore_dir is not None:
            restored = ckpt_utils.restore_latest(self.worker, restore_dir) is not None
        if restored:
            self.worker.mark_initialized()
        else:
            self.worker.initialize_variables(seed=seed, params=params)
        for ps in self.ps:
            ps.start()
        self.worker.attach_local_ps(self.ps)   # one-shot ps tasks are served by the worker after every launch
        self.worker.wait_ready()

    def close(self) -> None:
        self.worker.finish()
        for ps in self.ps:
            ps.stop()
        self.worker.close()
        for ps in self.ps:
            ps.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


@dataclass
class TrainLoopResult:
    steps_run: int
    last_global_step: int
    last_loss: float
    wall_s: float
    checkpoint_path: Optional[str] = None


def train_loop(worker: Worker, dataset: Dataset, train_steps: int = 4000, log_every: int = 100,
               checkpoint_dir: Optional[str] = None, save_checkpoint_secs: float = 600.0, seed: int = 0,
               chunk: int = 50, print_fn: Callable[[str], None] = print,
               inject_fault_after: int = 0, metrics: Optional[TrainMetricsWriter] = None,
               chunk_sleep_s: float = 0.0) -> TrainLoopResult:
    """The worker's `MonitoredTrainingSession` loop (DS:106-116).

    Runs until a step reports `global_step >= train_steps` (StopAtStepHook on the shared counter), printing
    `Train step {step}, loss: {loss}` whenever the step returned to *this* worker is a multiple of
    `log_every` (DS:115-116). The chief checkpoints every `save_checkpoint_secs` and at the end into
    `checkpoint_dir` (default: a fresh `tempfile.mkdtemp()`, DS:106).
    """
    loader = worker.make_loader(dataset.images, dataset.labels, seed=seed + 1000 * worker.task_index)
    if worker.is_chief and checkpoint_dir is None:
        checkpoint_dir = tempfile.mkdtemp()
    t0 = time.time()
    last_save = t0
    steps_run = 0
    last: Optional[StepOutput] = None
    ckpt_path = None
    stop = False
    while not stop:
        worker.heartbeat()     # liveness mark for the ps-side failure detector (--worker_timeout)
        with nvtx_range("dm.train.chunk"):
            outs = worker.run_steps(chunk, loader, stop_at_global_step=train_steps)
        steps_run += len(outs)
        if inject_fault_after and steps_run >= inject_fault_after:
            # fault injection (tests): die like a crashed process — no finish(), no close(), no atexit handlers
            print(f"[fault injection] worker {worker.task_index} dies after {steps_run} steps", flush=True)
            os._exit(42)
        if metrics is not None:
            metrics.update(outs)
        for o in outs:
            last = o
            if log_every and o.global_step % log_every == 0:
                print_fn("Train step {}, loss: {}".format(o.global_step, o.loss))
            if o.global_step >= train_steps:
                stop = True
        if not outs:
            break
        if chunk_sleep_s > 0:
            time.sleep(chunk_sleep_s)
        if worker.is_chief and checkpoint_dir and time.time() - last_save >= save_checkpoint_secs:
            with nvtx_range("dm.train.checkpoint"):
                ckpt_path = ckpt_utils.save_checkpoint(worker, checkpoint_dir)
            last_save = time.time()
    if worker.is_chief and checkpoint_dir:
        worker.drain()
        ckpt_path = ckpt_utils.save_checkpoint(worker, checkpoint_dir)
    return TrainLoopResult(steps_run, last.global_step if last else 0, last.loss if last else float("nan"),
                           time.time() - t0, ckpt_path)
