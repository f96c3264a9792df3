#!/bin/bash
# first GPU call of round 2: smoke (both engines, one-shot ps), fused trajectories, pipelined, bench N=1
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
python - > $O/probe.txt 2>&1 <<'PY'
try:
    from cuda import cuda
    cuda.cuInit(0)
    err, dev = cuda.cuDeviceGet(0)
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
                 "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_CLUSTER_LAUNCH"):
        a = getattr(cuda.CUdevice_attribute, name)
        print(name, cuda.cuDeviceGetAttribute(a, dev))
except Exception as e:
    print("probe failed", repr(e))
PY
timeout 240 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
for i in 0 1 2 3 9 10 11; do timeout 120 python bench_tools/gpu_e2e.py traj:$i > $O/traj_$i.log 2>&1; echo "rc=$?" >> $O/traj_$i.log; done
for i in 6 7 8 9 10; do timeout 150 python bench_tools/gpu_e2e.py pipelined:$i > $O/pipe_$i.log 2>&1; echo "rc=$?" >> $O/pipe_$i.log; done
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "rc=$?" >> $O/bench_k20.err
DM_PS_STATS=1 timeout 200 python bench.py --steps 2000 --warmup 50 > $O/bench_k2000.json 2> $O/bench_k2000.err; echo "rc=$?" >> $O/bench_k2000.err
tail -n 3 $O/smoke.log; for f in $O/traj_*.log $O/pipe_*.log; do tail -n 2 $f | head -c 600; done; cat $O/bench_k20.json | head -c 3000; tail -n 3 $O/bench_k20.err; cat $O/bench_k2000.json | head -c 1500; tail -n 5 $O/bench_k2000.err
