"""Toy parameter-server round trip — the equivalent of the reference's `ps_server-basic.py`.

The reference hard-codes a cluster (2 workers, 1 ps; PSB:22-35), selects the role by editing `isps` in the source
(PSB:37), hosts two 2x2 variables `w = 2`, `b = 5` on the ps (PSB:51-53) and has the worker print
`[w + b, w * b, w / b]` forever (PSB:62-63)  ->  7 / 10 / 0.4.

Here the role is a flag; the ps task exports the variables as a peer-memory segment (CUDA IPC on a GPU box, POSIX
shm otherwise), the worker maps it and evaluates the three expressions on *its* device straight from the ps's
memory. `--iterations 0` reproduces the endless loop.

    python examples/ps_server_basic.py --role ps &
    python examples/ps_server_basic.py --role worker --iterations 3
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dist_mnist_b200.cluster import ClusterSpec, Rendezvous  # noqa: E402
from dist_mnist_b200.parallel.peer_mem import Carver, Segment  # noqa: E402

CLUSTER = ClusterSpec(ps=("127.0.0.1:9910",), worker=("127.0.0.1:9900", "127.0.0.1:9901"))  # PSB:22-35


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--role", choices=["ps", "worker"], default="ps")   # reference: edit `isps` (PSB:37)
    ap.add_argument("--task_index", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=3, help="0 = loop forever like the reference")
    ap.add_argument("--backend", choices=["auto", "cuda", "cpu"], default="auto")
    ap.add_argument("--port", type=int, default=9910)
    args = ap.parse_args()
    cluster = ClusterSpec(ps=(f"127.0.0.1:{args.port}",), worker=CLUSTER.worker)
    backend = args.backend if args.backend != "auto" else ("cuda" if torch.cuda.is_available() else "cpu")
    kind = "cuda" if backend == "cuda" else "shm"
    rdv = Rendezvous(cluster, args.role, args.task_index, timeout_s=60)

    if args.role == "ps":
        c = Carver()
        c.add("w", 16)
        c.add("b", 16)
        seg = Segment.create(kind, c.total, device=0, table=c.table(), tag="toy")
        seg.tensor("w", torch.float32).fill_(2.0)    # constant_initializer(2), PSB:52
        seg.tensor("b", torch.float32).fill_(5.0)    # constant_initializer(5), PSB:53
        if kind == "cuda":
            torch.cuda.synchronize()
        rdv.put("toy/segment", seg.export())
        print("ps: serving w, b", flush=True)
        try:
            while rdv.try_get("toy/shutdown") is None:   # server.join() (PSB:45)
                time.sleep(0.1)
        except KeyboardInterrupt:
            pass
        seg.close()
        return 0

    dev_index = (1 + args.task_index) % max(1, torch.cuda.device_count()) if kind == "cuda" else -1
    if kind == "cuda":
        torch.cuda.set_device(dev_index)
    seg = Segment.open(rdv.get("toy/segment"), device=dev_index)
    i = 0
    while args.iterations == 0 or i < args.iterations:
        if kind == "cuda":
            # pull the ps-resident variables over NVLink (P2P copy) and compute on this worker's GPU
            w = torch.empty(4, device=f"cuda:{dev_index}")
            b = torch.empty(4, device=f"cuda:{dev_index}")
            from dist_mnist_b200 import _native as N
            N.check(N.lib().dm_memcpy_async(w.data_ptr(), seg.addr("w"), 16, None))
            N.check(N.lib().dm_memcpy_async(b.data_ptr(), seg.addr("b"), 16, None))
            torch.cuda.synchronize()
        else:
            w = seg.tensor("w", torch.float32).clone()
            b = seg.tensor("b", torch.float32).clone()
        w, b = w.view(2, 2), b.view(2, 2)
        print([(w + b).cpu().tolist(), (w * b).cpu().tolist(), (w / b).cpu().tolist()], flush=True)   # PSB:63
        i += 1
    if args.task_index == 0:
        rdv.put("toy/shutdown", True)
    seg.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
