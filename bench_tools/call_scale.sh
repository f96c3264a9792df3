#!/bin/bash
# usage: call_scale.sh N   — bench at N GPUs, driver config (K=20) twice + long (K=2000)
N=$1
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/scale_n$N; mkdir -p $O
P=$((29000 + RANDOM % 300))
for tag in k20 k20b; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "rc=$?" >> $O/bench_$tag.err
  P=$((P+701))
done
DM_PS_STATS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 2000 --warmup 50 > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "N=$N value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")}, d["config"]["barrier_sync_ms"], d["roofline"]["nvlink_frac"])
    except Exception as e: print(f, "ERR", e)
PY
for f in $O/bench_*.err; do grep -a "ps_stats\|rc=\|Error\|error" $f | cut -c1-500 | head -6; done
