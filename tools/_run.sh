set -u
python -m pytest tests/test_rolling_gpu.py -x -q -m gpu > gpurun_out/r2i_pytest_rolling.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2i_pytest_rolling.log
python tools/bench_ops.py rolling --reps 7 > gpurun_out/r2i_bench_ops_rolling.jsonl 2> gpurun_out/r2i_bench_ops.err; echo "ops rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2i_bench_ops_rolling.jsonl'):
    d=json.loads(l); print(d['case'], round(d['ms_median'],3), round(d['frac_of_measured_peak'],3))
PY
