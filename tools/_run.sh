python -m pytest tests/test_join_gpu.py -x -q -m gpu > gpurun_out/r2b_pytest_join.log 2>&1; tail -3 gpurun_out/r2b_pytest_join.log
python tools/time_join_c3.py > gpurun_out/r2b_time_join_c3.log 2>&1; cat gpurun_out/r2b_time_join_c3.log
python tools/time_groupby_async.py > gpurun_out/r2b_time_groupby_async.log 2>&1; cat gpurun_out/r2b_time_groupby_async.log
