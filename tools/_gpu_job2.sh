mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py -m gpu -q -x 2>&1 | tail -8
python tools/scan_bench.py 8191 4095 255x 67108867 2>&1 | cut -c1-150
for v in -4 -12 -16; do echo "== stages $((-v))"; python tools/scan_bench.py --variant=$v 8191x8190:0 4095x4097:0 2>&1 | grep -v "^#" | cut -c1-150; done
