cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu_v13.log
export DM_PDL=1
for cfg in "8 4" "16 4"; do set -- $cfg
  timeout 150 python bench.py --steps 8000 --warmup 50 --lanes $1 --graph_steps $2 2>&1 | grep "^{" > gpurun_out/bench_n1_v13_lanes$1_u$2.json
done
DM_GATHER_THREADS=1 timeout 150 python bench.py --steps 8000 --warmup 50 --lanes 8 --graph_steps 4 2>&1 | grep "^{" > gpurun_out/bench_n1_v13_lanes8_u4_gather1.json
DM_GATHER_THREADS=4 timeout 150 python bench.py --steps 8000 --warmup 50 --lanes 8 --graph_steps 4 2>&1 | grep "^{" > gpurun_out/bench_n1_v13_lanes8_u4_gather4.json
timeout 150 python bench.py --steps 8000 --warmup 50 --lanes 8 --graph_steps 4 --optimizer sgd --push_mode atomic --learning_rate 0.01 2>&1 | grep "^{" > gpurun_out/bench_n1_v13_sgd_atomic_lanes8_u4.json
nproc > gpurun_out/nproc.txt
for f in gpurun_out/bench_n1_v13_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'host us', d['config']['host_enqueue_us_per_step'], 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'])"; done
cat gpurun_out/nproc.txt gpurun_out/pytest_gpu_v13.log | cut -c1-300
