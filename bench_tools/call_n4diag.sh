#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/n4diag2; mkdir -p $O
timeout 60 python bench_tools/gpu_e2e.py traj:0 > $O/traj_0.log 2>&1; tail -n 1 $O/traj_0.log | cut -c1-250
timeout 60 python bench_tools/gpu_e2e.py traj:9 > $O/traj_9.log 2>&1; tail -n 1 $O/traj_9.log | cut -c1-250
P=$((29000 + RANDOM % 300))
DM_FUSED_DEBUG_TS=1 DM_PS_STATS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 4000 --warmup 50 --skip_e2e --skip_parity > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
P=$((P+701))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
for f in bench bench_k20; do python - <<PY
import json
d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f value", round(d["value"]), d["ms_per_step"], d.get("parity",{}).get("value_device_timed"))
PY
done
grep -a "step [0-9]:\|ps_stats\|rc=" $O/bench.err | cut -c1-560 | head -14
