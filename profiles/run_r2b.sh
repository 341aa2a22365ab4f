mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final2_pytest.log 2>&1; tail -4 gpurun_out/r2_final2_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
