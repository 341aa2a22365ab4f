mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_misc.py tests/test_gpu_lookup.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -12 gpurun_out/r2i_pytest.log
