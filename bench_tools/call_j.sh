#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_j; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
if grep -q "^OK" $O/hang_default.log; then
  for i in 7 8 9; do timeout 70 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; done
  for i in 0 2 10; do timeout 60 python bench_tools/gpu_e2e.py traj:$i > $O/traj_$i.log 2>&1; done
  DM_FUSED_DEBUG_TS=1 timeout 100 python bench.py --steps 2000 --warmup 50 > $O/bench_n1_k2000.json 2> $O/bench_n1_k2000.err
  timeout 100 python bench.py --steps 20 --warmup 5 > $O/bench_n1_k20.json 2> $O/bench_n1_k20.err
  timeout 100 python bench.py --steps 20 --warmup 5 > $O/bench_n1_k20b.json 2> $O/bench_n1_k20b.err
fi
tail -n 2 $O/hang_default.log; for f in $O/pipe_*.log $O/traj_*.log; do tail -n 1 $f | cut -c1-260; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/call_j/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")}, d["config"]["barrier_sync_ms"])
    except Exception as e: print(f, "ERR", e)
PY
grep -a "step [0-9]:" $O/bench_n1_k2000.err | cut -c1-600 | head -4
