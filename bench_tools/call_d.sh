#!/bin/bash
# 2-GPU call: bench N=2 (driver config + long), multi-GPU pytest
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_d; mkdir -p $O
P=$((29000 + RANDOM % 500))
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_k20.json 2> $O/bench_n2_k20.err; echo "rc=$?" >> $O/bench_n2_k20.err
P=$((P+700))
DM_PS_STATS=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2000 --warmup 50 > $O/bench_n2_k2000.json 2> $O/bench_n2_k2000.err; echo "rc=$?" >> $O/bench_n2_k2000.err
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
grep "^{" $O/bench_n2_k20.json | head -c 3000; tail -n 4 $O/bench_n2_k20.err | cut -c1-600; grep "^{" $O/bench_n2_k2000.json | head -c 3000; grep -a "ps_stats\|rc=" $O/bench_n2_k2000.err | cut -c1-600; tail -n 15 $O/pytest_multi.log | cut -c1-300
