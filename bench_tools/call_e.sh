#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_e; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
if grep -q "^OK" $O/hang_default.log; then
  for i in 0 2 9 11; do timeout 60 python bench_tools/gpu_e2e.py traj:$i > $O/traj_$i.log 2>&1; echo "rc=$?" >> $O/traj_$i.log; done
  for i in 7 8 9; do timeout 70 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; echo "rc=$?" >> $O/pipe_$i.log; done
  timeout 100 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
  DM_PS_STATS=1 timeout 120 python bench.py --steps 2000 --warmup 50 > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
  DM_FUSED_DEBUG_TS=1 timeout 60 python bench_tools/fused_phases.py > $O/phases.log 2>&1
fi
for f in $O/hang*.log; do echo "== $f"; tail -n 6 $f | cut -c1-300; done
for f in $O/traj_*.log $O/pipe_*.log; do tail -n 2 $f | head -c 500; done; python - <<'PY'
import json
for f in ("bench_k20","bench_k2000"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/call_e/{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", {k: round(v) for k,v in d["parity"].items() if k.startswith("value")})
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/bench_k20.err; grep -a "ps_stats" $O/bench_k2000.err | cut -c1-500; cat $O/phases.log | tail -n 14
