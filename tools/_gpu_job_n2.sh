mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"
tail -3 gpurun_out/r02_dist_check_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/r02_bench_n$N.err; cut -c1-300 gpurun_out/r02_bench_n$N.json
python -m pytest tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -3
