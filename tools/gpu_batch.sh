python -m pytest tests/test_groupby_gpu.py tests/test_strings_gpu.py tests/test_frame_gpu.py tests/test_numba_ext_gpu.py -q > gpurun_out/s5_pytest9.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s5_pytest9.log | cut -c1-220
python tools/ab_r1.py groupby > gpurun_out/s5_ab_gb6.jsonl 2> gpurun_out/s5_ab_gb6.err; echo "gb lean rc=$?"
SDCB200_GB_LEAN=0 python tools/ab_r1.py groupby > gpurun_out/s5_ab_gb6_off.jsonl 2> gpurun_out/s5_ab_gb6_off.err; echo "gb generic rc=$?"
