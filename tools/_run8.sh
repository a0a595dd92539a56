set -u
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err; echo "bench$N rc=$?"
if [ "$N" = "8" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/dist_check.py > gpurun_out/r2n_dist_check_n8.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/r2n_dist_check_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref_n$N.json 2> gpurun_out/r2n_bench_ref_n$N.err; echo "benchref$N rc=$?"
fi
