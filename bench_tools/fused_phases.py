"""In-kernel phase stamps of the fused step kernel (DM_FUSED_DEBUG_TS=1): clock64() of cluster 0 / CTA 0 at
step start, forward accumulator ready, after the reduce-scatter barrier, head done, after the all-gather barrier,
dW accumulator ready, tile pushed, step end — for the first 8 steps of the last launch. lanes=1 and lanes=8."""
import os
import sys

import torch

from dist_mnist_b200.models import mlp
from dist_mnist_b200.parallel.config import EngineConfig, OptimizerConfig
from dist_mnist_b200.session import InProcessCluster
from dist_mnist_b200.utils import data

NAMES = ["start", "acc1", "rs_bar", "head", "ag_bar", "acc2", "pushed", "end"]


def main():
    os.environ["DM_FUSED_DEBUG_TS"] = "1"
    ds = data.synthetic_mnist(8192, seed=0)
    spec = mlp.book_model(100)
    opt = OptimizerConfig("adam", 1e-4)
    for lanes in (1, 8):
        cfg = EngineConfig(backend="cuda", lanes=lanes, nslots=2 * lanes, pipeline_slots=max(4, 2 * lanes))
        with InProcessCluster(spec, opt, cfg, batch_size=32) as cl:
            w = cl.worker
            dev_x = ds.images.cuda().contiguous()
            dev_y = ds.labels.cuda().contiguous()
            n_rows = dev_x.shape[0] - 32
            for rep in range(3):
                w.run_resident(64, dev_x.data_ptr(), dev_y.data_ptr(), 784 * 4, 40, n_rows, rep * 64)
                w.wait_applied()
            ts = w._fx_dbg.cpu().view(8, 8)
            t0 = int(ts[0, 0])
            print(f"lanes={lanes}: cluster 0, cycles since its first step start (1965 MHz -> /1965 = us)")
            for j in range(8):
                row = [int(ts[j, k]) - t0 for k in range(8)]
                d = [row[k] - row[k - 1] if k else 0 for k in range(8)]
                print(f"  step {j}: " + " ".join(f"{NAMES[k]}=+{d[k]}" for k in range(1, 8)) +
                      f"  | step total {row[7] - row[0]} cyc = {(row[7] - row[0]) / 1965:.2f} us; start at {row[0]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
