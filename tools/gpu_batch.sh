python -m pytest tests/test_strings_gpu.py tests/test_frame_gpu.py -q > gpurun_out/s5_pytest13.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s5_pytest13.log | cut -c1-220
python tools/bench_ops.py strings --reps 5 2>gpurun_out/s5_str.err | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print(j['case'], round(j['ms_median'],3), round(j['frac_of_measured_peak'],3))"
