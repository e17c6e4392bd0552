mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_h.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_h.json 2> gpurun_out/r02_bench_n1_h.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_h.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_h.log; cut -c1-330 gpurun_out/r02_bench_n1_h.json; tail -2 gpurun_out/r02_bench_n1_h.err; tail -1 gpurun_out/r02_smoke_h.log
