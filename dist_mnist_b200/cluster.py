"""Cluster description, role/device mapping and the TCP rendezvous.

Reference parity (`/root/reference/distributed_server-basic.py`):
  * DS:71-78  `ClusterSpec({"worker": worker_hosts, "ps": ps_hosts})` from comma separated `host:port` lists.
  * DS:80     `tf.train.Server(cluster, job_name, task_index)` — one endpoint per task.
  * DS:108    chief election: `is_chief = (task_index == 0)`.

The new build keeps the `--ps_hosts/--worker_hosts` contract but uses the endpoints only for the control
plane: the first ps endpoint hosts a `torch.distributed.TCPStore` through which tasks exchange peer-memory
descriptors (CUDA IPC handles on GPUs, POSIX shm names on CPU), publish "variables initialised" / "serving"
marks and coordinate shutdown. The data plane never touches TCP.
"""
from __future__ import annotations

import datetime
import json
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

from torch.distributed import TCPStore


def parse_hosts(value: Optional[str], what: str) -> List[str]:
    """`"h1:p1,h2:p2"` -> `["h1:p1", "h2:p2"]` (reference DS:71,73 `.split(',')`)."""
    if value is None:
        raise ValueError(f"--{what} must be given (comma separated host:port list)")
    hosts = [h.strip() for h in value.split(",") if h.strip()]
    if not hosts:
        raise ValueError(f"--{what} is empty")
    for h in hosts:
        host, sep, port = h.rpartition(":")
        if not sep or not host or not port.isdigit():
            raise ValueError(f"--{what}: bad endpoint {h!r} (expected host:port)")
    return hosts


def split_endpoint(endpoint: str) -> Tuple[str, int]:
    host, _, port = endpoint.rpartition(":")
    return host, int(port)


@dataclass(frozen=True)
class ClusterSpec:
    """Static membership, exactly like the reference (no elastic join/leave)."""
    ps: Tuple[str, ...]
    worker: Tuple[str, ...]

    @staticmethod
    def from_flags(ps_hosts: Optional[str], worker_hosts: Optional[str]) -> "ClusterSpec":
        return ClusterSpec(tuple(parse_hosts(ps_hosts, "ps_hosts")), tuple(parse_hosts(worker_hosts, "worker_hosts")))

    @property
    def num_ps(self) -> int:
        return len(self.ps)

    @property
    def num_workers(self) -> int:
        return len(self.worker)

    @property
    def num_tasks(self) -> int:
        return self.num_ps + self.num_workers

    def as_dict(self) -> Dict[str, List[str]]:
        return {"worker": list(self.worker), "ps": list(self.ps)}

    def task_endpoint(self, job_name: str, task_index: int) -> str:
        jobs = {"ps": self.ps, "worker": self.worker}
        if job_name not in jobs:
            raise ValueError(f"unknown job name {job_name!r} (expected 'ps' or 'worker')")
        if not 0 <= task_index < len(jobs[job_name]):
            raise ValueError(f"task index {task_index} out of range for job {job_name!r} ({len(jobs[job_name])} tasks)")
        return jobs[job_name][task_index]

    def global_rank(self, job_name: str, task_index: int) -> int:
        """ps tasks first, then workers — used for default GPU assignment."""
        self.task_endpoint(job_name, task_index)
        return task_index if job_name == "ps" else self.num_ps + task_index

    def rendezvous_endpoint(self) -> Tuple[str, int]:
        return split_endpoint(self.ps[0])


def default_device_index(cluster: ClusterSpec, job_name: str, task_index: int, n_devices: int,
                         colocate: bool = False) -> int:
    """GPU for a task on a single box: ps k -> GPU k, worker i -> GPU (num_ps + i), wrapping around.

    With fewer GPUs than tasks (e.g. everything on one GPU) tasks share devices; `colocate=True` maps
    worker i onto GPU i so that ps 0 and worker 0 share GPU 0 (N workers on N GPUs).
    """
    if n_devices <= 0:
        return -1
    if colocate and job_name == "worker":
        return task_index % n_devices
    return cluster.global_rank(job_name, task_index) % n_devices


class Rendezvous:
    """Thin JSON key-value layer over TCPStore hosted by ps task 0."""

    def __init__(self, cluster: ClusterSpec, job_name: str, task_index: int, timeout_s: float = 300.0,
                 endpoint: Optional[Tuple[str, int]] = None):
        self.cluster = cluster
        self.is_master = job_name == "ps" and task_index == 0
        host, port = endpoint if endpoint is not None else cluster.rendezvous_endpoint()
        self._endpoint = (host, port)
        self.timeout_s = timeout_s
        last_err: Optional[Exception] = None
        deadline = time.time() + timeout_s
        self.store = None
        while time.time() < deadline:
            try:
                self.store = TCPStore(host, port, world_size=None, is_master=self.is_master,
                                      timeout=datetime.timedelta(seconds=timeout_s), wait_for_workers=False)
                break
            except Exception as e:  # the master may not be up yet (tasks start in any order)
                last_err = e
                if self.is_master:
                    raise
                time.sleep(0.2)
        if self.store is None:
            raise TimeoutError(f"could not reach the rendezvous store at {host}:{port}: {last_err}")

    def clone(self) -> "Rendezvous":
        """A second client connection to the same store (for a helper thread)."""
        c = Rendezvous.__new__(Rendezvous)
        c.cluster, c.is_master, c.timeout_s = self.cluster, False, self.timeout_s
        c._endpoint = self._endpoint
        host, port = self._endpoint
        c.store = TCPStore(host, port, world_size=None, is_master=False,
                           timeout=datetime.timedelta(seconds=self.timeout_s), wait_for_workers=False)
        return c

    def put(self, key: str, value) -> None:
        self.store.set(key, json.dumps(value))

    def get(self, key: str, timeout_s: Optional[float] = None):
        t = self.timeout_s if timeout_s is None else timeout_s
        self.store.wait([key], datetime.timedelta(seconds=t))
        return json.loads(self.store.get(key).decode())

    def try_get(self, key: str):
        try:
            if not self.store.check([key]):
                return None
        except Exception:
            return None
        return json.loads(self.store.get(key).decode())

    def add(self, key: str, amount: int = 1) -> int:
        return int(self.store.add(key, amount))

    def wait_count(self, key: str, target: int, timeout_s: Optional[float] = None, poll_s: float = 0.01) -> None:
        t = self.timeout_s if timeout_s is None else timeout_s
        deadline = time.time() + t
        while self.add(key, 0) < target:
            if time.time() > deadline:
                raise TimeoutError(f"rendezvous: {key} did not reach {target} within {t}s")
            time.sleep(poll_s)
