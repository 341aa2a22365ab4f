timeout 100 python -m pytest tests/test_gpu_update.py -q -x -k more_than_eight 2>&1 | tail -8
