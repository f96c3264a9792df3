cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/pytest_gpu_v18.log
timeout 100 python bench.py 2>&1 | grep "^{" > gpurun_out/bench_n1_v18_default.json
timeout 100 python bench.py --steps 500 --warmup 10 2>&1 | grep "^{" > gpurun_out/bench_n1_v18_k500.json
timeout 100 python bench.py --steps 20000 2>&1 | grep "^{" > gpurun_out/bench_n1_v18_k20000.json
for f in gpurun_out/bench_n1_v18_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'], d['gpu_launches'])"; done
cat gpurun_out/pytest_gpu_v18.log
