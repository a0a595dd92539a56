set -u
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_check.py > gpurun_out/r2j_dist_check_n2.log 2>&1; echo "dist_check rc=$?"; tail -12 gpurun_out/r2j_dist_check_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench2 rc=$?"
