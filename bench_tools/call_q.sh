#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_q; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
for i in 7 8; do timeout 70 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; done
for t in a b; do DM_FEXEC_TRACE=1 timeout 100 python bench.py --steps 20 --warmup 5 > $O/n1_k20_$t.json 2> $O/n1_k20_$t.err; done
timeout 100 python bench.py --steps 2000 --warmup 50 > $O/n1_k2000.json 2> $O/n1_k2000.err
tail -n 2 $O/hang_default.log; for f in $O/pipe_*.log; do tail -n 1 $f | cut -c1-260; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/call_q/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "us total", round(d["ms_per_step"]*d["steps"]*1e3,1), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")})
    except Exception as e: print(f, "ERR", e)
PY
grep -a "fexec" $O/n1_k20_a.err | head -24
