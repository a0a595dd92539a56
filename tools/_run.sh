set -u
python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python tools/bench_ops.py rolling groupby join --reps 5 > gpurun_out/r2m_bench_ops.jsonl 2> gpurun_out/r2m_bench_ops.err; echo "ops rc=$?"
