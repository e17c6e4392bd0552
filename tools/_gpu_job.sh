mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu_g.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err
python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_g.json 2> gpurun_out/r02_bench_ref_g.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_g.log 2>&1
python tools/scan_bench.py > gpurun_out/r02_scan_bench.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_g.log; cut -c1-400 gpurun_out/r02_bench_n1_g.json; tail -3 gpurun_out/r02_bench_n1_g.err; tail -2 gpurun_out/r02_smoke_g.log; cat gpurun_out/r02_scan_bench.log | awk '{print $1,$2,$4,$6,$7}'
