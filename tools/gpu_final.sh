#!/bin/bash
# Round-end refresh on one B200 (run under gpurun): full GPU test suite, the headline bench, the per-operator table,
# ncu launch lists of the changed operators and one --set full capture of each new join kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench ref rc=$?"
python tools/bench_ops.py rolling groupby join sort strings --reps 10 > gpurun_out/f_bench_ops.jsonl 2> gpurun_out/f_bench_ops.err; echo "ops rc=$?"
python tools/bench_dist.py --cases groupby_lo,join_bcast --reps 5 > gpurun_out/f_dist_n1.jsonl 2> gpurun_out/f_dist_n1.err; echo "dist n=1 rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/f_l_join.csv python tools/prof_case.py join 2 > gpurun_out/f_ncu_join.log 2>&1; echo "ncu join rc=$?"
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/f_l_gb100m.csv python tools/prof_case.py groupby_100m 2 > gpurun_out/f_ncu_gb.log 2>&1; echo "ncu gb rc=$?"
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/f_l_moments.csv python tools/prof_case.py rolling_skew 2 > gpurun_out/f_ncu_mom.log 2>&1; echo "ncu mom rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_unique_row1 -c 1 -f -o gpurun_out/f_join_row1 python tools/prof_case.py join 1 > gpurun_out/f_ncu_j1.log 2>&1; echo "ncu row1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_unique_compact -c 1 -f -o gpurun_out/f_join_compact python tools/prof_case.py join 1 > gpurun_out/f_ncu_j2.log 2>&1; echo "ncu compact rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gb_smem -c 1 -f -o gpurun_out/f_gb100m python tools/prof_case.py groupby_100m 1 > gpurun_out/f_ncu_g1.log 2>&1; echo "ncu gb full rc=$?"
