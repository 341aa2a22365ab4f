mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
for occ in 4 5 6; do
HB_SHORT_OCC=$occ timeout 120 python bench.py --steps 30 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('occ',$occ, l['ms_per_step'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})"
done 2>&1 | tee gpurun_out/r2b_occ.log
