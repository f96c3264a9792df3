#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/final1; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -n 14 $O/pytest_gpu.log | cut -c1-200
for t in a b; do timeout 100 python bench.py --steps 20 --warmup 5 > $O/n1_k20_$t.json 2> $O/n1_k20_$t.err; done
DM_PS_STATS=1 timeout 100 python bench.py --steps 2000 --warmup 50 > $O/n1_k2000.json 2> $O/n1_k2000.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final1/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")})
    except Exception as e: print(f, "ERR", e)
PY
grep -a "ps_stats" $O/n1_k2000.err | head -3 | cut -c1-400
bash bench_tools/make_evidence.sh > $O/evidence.log 2>&1; tail -n 12 $O/evidence.log | cut -c1-200
