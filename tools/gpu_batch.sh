python -m pytest tests/test_numba_ext_gpu.py tests/test_frame_gpu.py tests/test_rolling_moments_gpu.py tests/test_groupby_gpu.py tests/test_join_gpu.py -q > gpurun_out/s5_pytest8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s5_pytest8.log | cut -c1-220
python tools/bench_ops.py rolling --reps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l)
    if any(k in j['case'] for k in ('skew','kurt','cov','corr')): print(j['case'], round(j['ms_median'],3), round(j['frac_of_measured_peak'],3))"
