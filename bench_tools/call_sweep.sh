#!/bin/bash
N=$1
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
P=$((29000 + RANDOM % 300))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P -m bench_tools.p2p_sweep --max_bytes 1073741824 --out gpurun_out/p2p_sweep_n${N}_1g.json > gpurun_out/p2p_sweep_n${N}.log 2>&1; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/p2p_sweep_n${N}_1g.json"))
print("flag latency us", d["one_way_flag_latency_us"])
for r in d["rows"]:
    print(r["bytes"], "push_tma %.1f pull_tma %.1f push_ldst %.1f pull_ldst %.1f GB/s/worker | at ps push %.0f pull %.0f | nccl red %.1f bcast %.1f | host-staged %.1f" % (r["push_tma_GBps_per_worker"], r["pull_tma_GBps_per_worker"], r["push_ldst_GBps_per_worker"], r["pull_ldst_GBps_per_worker"], r["push_tma_GBps_at_ps"], r["pull_tma_GBps_at_ps"], r["nccl_reduce_GBps_per_worker"], r["nccl_broadcast_GBps_per_worker"], r["host_staged_standin_GBps_per_worker"]))
PY
