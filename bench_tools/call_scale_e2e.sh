#!/bin/bash
N=$1
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/scale_e2e_n$N; mkdir -p $O
(cat /sys/fs/cgroup/cpu.max; nproc; python -c "import os; print(len(os.sched_getaffinity(0)))") > $O/cpu.txt 2>&1; cat $O/cpu.txt
P=$((29000 + RANDOM % 300))
DM_FEXEC_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 --skip_parity > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
P=$((P+701))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 2000 --warmup 50 --skip_parity > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
P=$((P+701))
DM_GATHER_THREADS=4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 2000 --warmup 50 --skip_parity > $O/bench_k2000_t4.json 2> $O/bench_k2000_t4.err; echo "rc=$?" >> $O/bench_k2000_t4.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "N=$N value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), d["config"].get("usable_cores"), d["config"].get("host_cores"))
    except Exception as e: print(f, "ERR", e)
PY
grep -a "fexec" $O/bench_k20.err | tail -n 12
