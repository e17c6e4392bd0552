mkdir -p gpurun_out
python tools/scan_bench.py > gpurun_out/r02_scan_bench.log 2>&1
cat gpurun_out/r02_scan_bench.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/cub_scan_bar tools/cub_scan_bar.cu 2>/dev/null && /tmp/cub_scan_bar > gpurun_out/r02_cub_scan_bar.log 2>&1; cat gpurun_out/r02_cub_scan_bar.log
ITERS=1 WARM=1 timeout 900 ncu --set full --clock-control none -k regex:k_scan -f -o /tmp/r02_scan python tools/scan_bench.py 67108864:None:float32 33554432:None:float64 1048576x64:0 64x1048576:1 64x1048576:0 uint8 > gpurun_out/r02_ncu_scan.log 2>&1
python tools/ncu_summary.py /tmp/r02_scan.ncu-rep gpurun_out/r02_ncu_summary_scan.csv > gpurun_out/r02_ncu_summary_scan.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
tail -2 gpurun_out/r02_bench_n1_c.err
