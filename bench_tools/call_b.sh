#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_b; mkdir -p $O
timeout 120 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
DM_PS_STATS=1 timeout 150 python bench.py --steps 2000 --warmup 50 > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
for i in 0 1 3 9 10 11; do timeout 90 python bench_tools/gpu_e2e.py traj:$i > $O/traj_$i.log 2>&1; echo "rc=$?" >> $O/traj_$i.log; done
for i in 6 7 9 10; do timeout 100 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; echo "rc=$?" >> $O/pipe_$i.log; done
DM_FUSED_DEBUG_TS=1 timeout 100 python bench_tools/fused_phases.py > $O/phases.log 2>&1
for f in $O/traj_*.log $O/pipe_*.log; do tail -n 2 $f | head -c 500; done; cat $O/bench_k20.json | head -c 3500; tail -n 3 $O/bench_k20.err; cat $O/bench_k2000.json | head -c 3500; tail -n 5 $O/bench_k2000.err; cat $O/phases.log | tail -n 30
