#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_p; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
timeout 70 python bench_tools/gpu_e2e.py pipelined:7 > $O/pipe_7.log 2>&1
for t in a b; do timeout 100 python bench.py --steps 20 --warmup 5 > $O/n1_k20_$t.json 2> $O/n1_k20_$t.err; done
timeout 100 python bench.py --steps 2000 --warmup 50 > $O/n1_k2000.json 2> $O/n1_k2000.err
P=$((29000 + RANDOM % 300))
for t in a b; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $O/n2_k20_$t.json 2> $O/n2_k20_$t.err; P=$((P+701))
done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2000 --warmup 50 > $O/n2_k2000.json 2> $O/n2_k2000.err
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -n 3 $O/pytest_multi.log | cut -c1-300
tail -n 2 $O/hang_default.log; tail -n 1 $O/pipe_7.log | cut -c1-260
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/call_p/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "us total", round(d["ms_per_step"]*d["steps"]*1e3,1), "parity", {k: round(v) for k,v in d.get("parity",{}).items() if k.startswith("value")}, "enq us/step", d["config"]["host_enqueue_us_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
