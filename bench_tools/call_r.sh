#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_r; mkdir -p $O
timeout 300 python -m bench_tools.gpu_check > $O/gpu_check.log 2>&1; echo "rc=$?" >> $O/gpu_check.log
grep -a "FAIL\|rc=" $O/gpu_check.log | head -20 | cut -c1-300; grep -ac PASS $O/gpu_check.log
grep -a "fwd dt=0" $O/gpu_check.log | head -12 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q > $O/pytest.log 2>&1; tail -n 4 $O/pytest.log | cut -c1-300
