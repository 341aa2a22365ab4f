mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -6 gpurun_out/r2e_pytest.log
timeout 300 python bench.py > gpurun_out/r2e_bench_n1.log 2>&1; tail -1 gpurun_out/r2e_bench_n1.log | python -c "
import sys,json; l=json.loads(sys.stdin.read())
print(l['ms_per_step'], l['value'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})
print('e2e', l['e2e']['ms_per_step'], l['e2e_variants']['device_out']['ms_per_step'], 'uniform', l['extra']['uniform']['ms_per_step'], l['extra']['uniform']['kernels_ms'])
print([(r['kernel'], round(r['frac'],3)) for r in l['roofline_all']], l['hbm_roofline_note'])
print('cpu', l['cpu_baseline']['ms_per_step'])
"
timeout 200 python bench.py --dim 64 --no-cpu-baseline > gpurun_out/r2e_bench_n1_d64.log 2>&1; tail -1 gpurun_out/r2e_bench_n1_d64.log | python -c "
import sys,json; l=json.loads(sys.stdin.read())
print('D64', l['ms_per_step'], l['value'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})"
