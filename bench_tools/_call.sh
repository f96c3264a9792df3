cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu_v16.log
export DM_PS_STATS=1
timeout 100 python bench.py --steps 8000 --warmup 50 > gpurun_out/bench_n1_v16_full.log 2>&1
grep "^{" gpurun_out/bench_n1_v16_full.log > gpurun_out/bench_n1_v16_adam.json; grep "ps_stats" gpurun_out/bench_n1_v16_full.log
timeout 100 python bench.py --steps 8000 --warmup 50 --lanes 16 > gpurun_out/bench_n1_v16_full2.log 2>&1
grep "^{" gpurun_out/bench_n1_v16_full2.log > gpurun_out/bench_n1_v16_adam_lanes16.json; grep "ps_stats" gpurun_out/bench_n1_v16_full2.log
timeout 100 python bench.py --steps 4000 --warmup 50 --lanes 1 2>&1 | grep "^{" > gpurun_out/bench_n1_v16_adam_lanes1.json
unset DM_PS_STATS
timeout 120 python -m bench_tools.profile_kernels --graph_time --pdl > gpurun_out/kernel_times_v16.log 2>&1
for f in gpurun_out/bench_n*_v16_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'])"; done
cat gpurun_out/kernel_times_v16.log gpurun_out/pytest_gpu_v16.log | cut -c1-250
