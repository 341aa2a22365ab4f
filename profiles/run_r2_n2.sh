mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi.log 2>&1; tail -5 gpurun_out/r2_pytest_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_bench_n2.log 2>&1; tail -1 gpurun_out/r2_bench_n2.log | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})"
