python -m pytest tests/test_join_gpu.py tests/test_frame_gpu.py -x -q -m gpu > gpurun_out/r2k_pytest_join.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2k_pytest_join.log
python tools/time_join_c3.py > gpurun_out/r2k_time_join_c3.log 2>&1; cat gpurun_out/r2k_time_join_c3.log
