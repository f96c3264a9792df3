cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/pytest_gpu_v19_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
DM_PS_STATS=1 timeout 200 $TR --master-port 29911 bench.py --gpus 2 > gpurun_out/bench_n2_v19_full.log 2>&1
grep "^{" gpurun_out/bench_n2_v19_full.log > gpurun_out/bench_n2_v19_default.json; grep ps_stats gpurun_out/bench_n2_v19_full.log
timeout 100 python bench.py --impl reference
for f in gpurun_out/bench_n2_v19_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'], d['gpu_launches'])"; done
cat gpurun_out/pytest_gpu_v19_2gpu.log
