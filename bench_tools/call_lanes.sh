#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/lanes; mkdir -p $O
for L in 8 12 16; do
  timeout 100 python bench.py --steps 20 --warmup 5 --lanes $L --skip_parity > $O/n1_k20_l$L.json 2> $O/n1_k20_l$L.err
  timeout 100 python bench.py --steps 2000 --warmup 50 --lanes $L --skip_parity --skip_e2e > $O/n1_k2000_l$L.json 2> $O/n1_k2000_l$L.err
done
P=$((29000 + RANDOM % 300))
for L in 8 16; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --lanes $L --skip_parity > $O/n2_k20_l$L.json 2> $O/n2_k20_l$L.err
  P=$((P+701))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2000 --warmup 50 --lanes $L --skip_parity --skip_e2e > $O/n2_k2000_l$L.json 2> $O/n2_k2000_l$L.err
  P=$((P+701))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/lanes/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "us total", round(d["ms_per_step"]*d["steps"]*1e3,1))
    except Exception as e: print(f, "ERR", e)
PY
