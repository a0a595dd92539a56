set -u
python tools/sanitize_cases.py > gpurun_out/r2q_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r2q_plain.log
for f in rolling groupby join sort strings shuffle; do
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_cases.py $f > gpurun_out/r2q_memcheck_$f.log 2>&1; echo "memcheck $f rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/r2q_memcheck_$f.log; tail -2 gpurun_out/r2q_memcheck_$f.log
done
python tools/time_groupby_async.py > gpurun_out/r2q_time_groupby_async.log 2>&1; cat gpurun_out/r2q_time_groupby_async.log
