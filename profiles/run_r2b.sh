mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
timeout 300 python bench.py > gpurun_out/r2h_bench_n1.log 2>&1; tail -1 gpurun_out/r2h_bench_n1.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()}, l['e2e']['ms_per_step'])"
