#!/bin/bash
# Round-end refresh on one B200 (run under gpurun): full GPU test suite, smoke, the headline bench (both arms), the
# per-operator groupby table, ncu launch lists of the groupby cases and one --set full capture of the lean kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python tools/bench_ops.py groupby --reps 10 > gpurun_out/g_bench_ops.jsonl 2> gpurun_out/g_bench_ops.err; echo "ops rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/g_l_gb.csv python tools/prof_case.py groupby 3 > gpurun_out/g_ncu_gb.log 2>&1; echo "ncu gb rc=$?"
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/g_l_gb100m.csv python tools/prof_case.py groupby_100m 2 > gpurun_out/g_ncu_gb100m.log 2>&1; echo "ncu gb100m rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gb_smem -c 1 -f -o gpurun_out/g_gb100m python tools/prof_case.py groupby_100m 1 > gpurun_out/g_ncu_g1.log 2>&1; echo "ncu gb full rc=$?"
