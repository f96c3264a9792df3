#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/n4e2e; mkdir -p $O
P=$((29000 + RANDOM % 300))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
P=$((P+701))
DM_PS_STATS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 2000 --warmup 50 --skip_parity > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "N=4 value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), d["config"].get("usable_cores"), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")})
    except Exception as e: print(f, "ERR", e)
PY
grep -a "ps_stats sync 2\|ps_stats\] sync 2" $O/bench_k2000.err | cut -c1-400
