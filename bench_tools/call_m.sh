#!/bin/bash
cd /root/repo; mkdir -p gpurun_out; export PYTHONPATH=/root/repo
O=gpurun_out/call_m; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -s > $O/pytest_fused.log 2>&1; echo "rc=$?" >> $O/pytest_fused.log
grep -a "parity\|passed\|failed\|Error\|assert\|rc=" $O/pytest_fused.log | cut -c1-900 | tail -n 20
bash bench_tools/make_evidence.sh
