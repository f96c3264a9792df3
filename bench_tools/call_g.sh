#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_h; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
if grep -q "^OK" $O/hang_default.log; then
  timeout 60 python bench_tools/gpu_e2e.py pipelined:7 > $O/pipe_7.log 2>&1
  timeout 60 python bench_tools/gpu_e2e.py pipelined:10 > $O/pipe_10.log 2>&1
  timeout 60 python bench_tools/gpu_e2e.py traj:2 > $O/traj_2.log 2>&1
  DM_FUSED_DEBUG_TS=1 timeout 100 python bench.py --steps 2000 --warmup 50 > $O/bench_n1_k2000.json 2> $O/bench_n1_k2000.err
  P=$((29000 + RANDOM % 500))
  DM_FUSED_DEBUG_TS=1 DM_BENCH_PROBE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_k20.json 2> $O/bench_n2_k20.err; echo "rc=$?" >> $O/bench_n2_k20.err
  P=$((P+700))
  DM_FUSED_DEBUG_TS=1 DM_PS_STATS=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2000 --warmup 50 > $O/bench_n2_k2000.json 2> $O/bench_n2_k2000.err; echo "rc=$?" >> $O/bench_n2_k2000.err
fi
tail -n 3 $O/hang_default.log; for f in $O/pipe_7.log $O/pipe_10.log $O/traj_2.log; do tail -n 1 $f | cut -c1-300; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/call_h/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")})
    except Exception as e: print(f, "ERR", e)
PY
for f in $O/bench_*.err; do echo "== $f"; grep -a "step [0-9]:\|ps_stats\|rc=\|Error\|error\|probe" $f | cut -c1-700 | head -16; done
