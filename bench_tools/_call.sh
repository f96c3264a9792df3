cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
export DM_PS_STATS=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29911 bench.py --gpus 8 --steps 8000 --warmup 50 > gpurun_out/bench_n8_v16_full.log 2>&1
grep "^{" gpurun_out/bench_n8_v16_full.log > gpurun_out/bench_n8_v16_adam.json; grep "ps_stats" gpurun_out/bench_n8_v16_full.log; tail -2 gpurun_out/bench_n8_v16_full.log | cut -c1-200
timeout 240 $TR --master-port 29921 bench.py --gpus 8 --steps 8000 --warmup 50 --optimizer sgd --push_mode atomic --learning_rate 0.01 2>&1 | grep "^{" > gpurun_out/bench_n8_v16_sgd_atomic.json
timeout 240 $TR --master-port 29931 bench.py --gpus 8 --steps 3000 --warmup 50 --num_ps 2 --model wide --dtype bf16 --lanes 4 --graph_steps 2 > gpurun_out/bench_n8_v16_wide_full.log 2>&1
grep "^{" gpurun_out/bench_n8_v16_wide_full.log > gpurun_out/bench_n8_v16_wide_bf16_2ps.json; grep "ps_stats" gpurun_out/bench_n8_v16_wide_full.log
for f in gpurun_out/bench_n8_v16_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'])"; done
