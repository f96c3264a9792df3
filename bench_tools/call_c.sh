#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_c; mkdir -p $O
timeout 60 python bench_tools/debug_hang.py 0 fused > $O/hang_default.log 2>&1; echo "rc=$?" >> $O/hang_default.log
if grep -q "^OK" $O/hang_default.log; then
  timeout 100 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
  DM_PS_STATS=1 timeout 120 python bench.py --steps 2000 --warmup 50 > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
  for i in 0 1 3 9 10 11; do timeout 60 python bench_tools/gpu_e2e.py traj:$i > $O/traj_$i.log 2>&1; echo "rc=$?" >> $O/traj_$i.log; done
  for i in 6 7 9 10; do timeout 70 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; echo "rc=$?" >> $O/pipe_$i.log; done
  DM_FUSED_DEBUG_TS=1 timeout 60 python bench_tools/fused_phases.py > $O/phases.log 2>&1
else
  timeout 60 python bench_tools/debug_hang.py 8 fused > $O/hang_ctas8.log 2>&1; echo "rc=$?" >> $O/hang_ctas8.log
  timeout 60 python bench_tools/debug_hang.py 0 graph > $O/hang_graph.log 2>&1; echo "rc=$?" >> $O/hang_graph.log
fi
for f in $O/hang*.log; do echo "== $f"; tail -n 25 $f | cut -c1-400; done
for f in $O/traj_*.log $O/pipe_*.log; do tail -n 2 $f | head -c 500; done; cat $O/bench_k20.json | head -c 3500; tail -n 3 $O/bench_k20.err; cat $O/bench_k2000.json | head -c 3500; tail -n 5 $O/bench_k2000.err; cat $O/phases.log | tail -n 30
