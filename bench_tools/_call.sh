cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
echo "== graph times no pdl" > gpurun_out/kernel_times_v8.log
timeout 120 python bench_tools/profile_kernels.py --graph_time >> gpurun_out/kernel_times_v8.log 2>&1
echo "== graph times pdl" >> gpurun_out/kernel_times_v8.log
timeout 120 python bench_tools/profile_kernels.py --graph_time --pdl >> gpurun_out/kernel_times_v8.log 2>&1
for cfg in "1 0 2" "2 0 2" "2 1 2" "4 1 4" "4 0 4"; do set -- $cfg
  DM_PDL=$2 timeout 150 python bench.py --steps 4000 --warmup 50 --lanes $1 --nslots $3 2>&1 | grep "^{" > gpurun_out/bench_n1_v9_lanes$1_pdl$2.json
done
DM_PDL=1 timeout 150 python bench.py --steps 4000 --warmup 50 --lanes 2 --optimizer sgd --push_mode atomic --learning_rate 0.01 2>&1 | grep "^{" > gpurun_out/bench_n1_v9_sgd_atomic_lanes2_pdl1.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"head_kernel|gemm_tcgen05" -s 3 -c 3 -f -o gpurun_out/step_v8 python bench_tools/profile_kernels.py --iters 3 > gpurun_out/ncu_v8_stdout.txt 2>&1
tail -3 gpurun_out/ncu_v8_stdout.txt
for f in gpurun_out/bench_n1_v9_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'host us', d['config']['host_enqueue_us_per_step'], 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
cat gpurun_out/kernel_times_v8.log
