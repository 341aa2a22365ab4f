mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_local_group.py tests/test_gpu_update.py tests/test_gpu_misc.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -5 gpurun_out/r2g_pytest.log
