cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
(timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" 2>&1 | tail -3) > gpurun_out/smoke_v17.log
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest_gpu_v17.log
DM_PS_STATS=1 timeout 100 python bench.py > gpurun_out/bench_n1_v17_full.log 2>&1
grep "^{" gpurun_out/bench_n1_v17_full.log > gpurun_out/bench_n1_v17_default.json; grep "ps_stats" gpurun_out/bench_n1_v17_full.log
timeout 100 python bench.py --steps 8000 --graph_steps 8 2>&1 | grep "^{" > gpurun_out/bench_n1_v17_u8.json
DM_GATHER_THREADS=6 timeout 100 python bench.py --steps 8000 2>&1 | grep "^{" > gpurun_out/bench_n1_v17_gather6.json
(timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python -m bench_tools.profile_kernels --iters 2 2>&1 | tail -12) > gpurun_out/sanitizer_memcheck_v17.log
for f in gpurun_out/bench_n1_v17_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(' value', round(d['value']), 'us/step', round(d['ms_per_step']*1e3,2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['config']['global_step_after_run'], d['gpu_launches'])"; done
cat gpurun_out/smoke_v17.log gpurun_out/pytest_gpu_v17.log gpurun_out/sanitizer_memcheck_v17.log | cut -c1-250
