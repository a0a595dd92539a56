set -u
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python bench.py --steps 2 --warmup 1 --no-cpu --no-ops > gpurun_out/r2p_bench_short.json 2> gpurun_out/r2p_bench_short.err; echo "bench rc=$?"
ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2p_l_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ops > gpurun_out/r2p_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
for c in rolling rolling_std rolling_mean_nan; do
ncu --set full --clock-control none --import-source on -k regex:rolling_run -c 1 -f -o gpurun_out/r2p_$c python tools/prof_case.py $c 1 > gpurun_out/r2p_ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
ncu --metrics $M --clock-control none -c 80 --csv --log-file gpurun_out/r2p_l_join.csv python tools/prof_case.py join 2 > gpurun_out/r2p_ncu_join.log 2>&1; echo "ncu join rc=$?"
ncu --metrics $M --clock-control none -c 80 --csv --log-file gpurun_out/r2p_l_groupby.csv python tools/prof_case.py groupby 3 > gpurun_out/r2p_ncu_gb.log 2>&1; echo "ncu gb rc=$?"
