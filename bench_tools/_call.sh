cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
export DM_FUSED_HEAD=1
export DM_TRAJ_VERBOSE=
(for i in 0 1 2 3 8; do timeout 60 python bench_tools/gpu_e2e.py traj:$i 2>&1 | tail -2; done) > gpurun_out/fused_traj.log 2>&1
(timeout 90 python bench_tools/gpu_e2e.py pipelined:1 2>&1 | tail -2; timeout 90 python bench_tools/gpu_e2e.py pipelined:2 2>&1 | tail -2) > gpurun_out/fused_pipe.log 2>&1
timeout 60 python -m bench_tools.profile_kernels --graph_time --pdl --fuse_head > gpurun_out/kernel_times_fused.log 2>&1
timeout 100 python bench.py --steps 8000 2>&1 | grep "^{" > gpurun_out/bench_n1_fused.json
timeout 100 python bench.py --steps 4000 --lanes 1 2>&1 | grep "^{" > gpurun_out/bench_n1_fused_lanes1.json
for f in gpurun_out/bench_n1_fused*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['config']['global_step_after_run'], d['gpu_launches'], d['kernels_per_step'])"; done
cat gpurun_out/fused_traj.log gpurun_out/fused_pipe.log gpurun_out/kernel_times_fused.log | cut -c1-260
