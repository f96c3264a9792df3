#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_o; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
if grep -q "^OK" $O/hang_default.log; then
  for i in 7 8 10; do timeout 70 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; done
  timeout 200 python -m pytest tests/test_gpu_fused.py -x -q > $O/pytest_fused.log 2>&1
  for L in 8 12; do
    DM_FUSED_DEBUG_TS=1 timeout 100 python bench.py --steps 20 --warmup 5 --lanes $L > $O/n1_k20_l$L.json 2> $O/n1_k20_l$L.err
  done
  timeout 100 python bench.py --steps 2000 --warmup 50 --lanes 12 > $O/n1_k2000_l12.json 2> $O/n1_k2000_l12.err
fi
tail -n 2 $O/hang_default.log; for f in $O/pipe_*.log; do tail -n 1 $f | cut -c1-260; done; tail -n 3 $O/pytest_fused.log | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/call_o/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "us total", round(d["ms_per_step"]*d["steps"]*1e3,1), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")}, "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
grep -a "globaltimer\|step [0-3]:" $O/n1_k20_l8.err | cut -c1-420
