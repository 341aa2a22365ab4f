mkdir -p gpurun_out
python tools/sort_timing.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_local_group.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
timeout 120 python bench.py --steps 30 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})" 2>&1 | tee gpurun_out/r2d_bench.log
