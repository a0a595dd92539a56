python -m pytest tests/test_rolling_gpu.py -x -q -m gpu > gpurun_out/r2b_pytest_rolling.log 2>&1; tail -15 gpurun_out/r2b_pytest_rolling.log
