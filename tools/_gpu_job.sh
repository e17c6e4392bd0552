timeout 300 python tools/scan_bench.py 67108864:None:float32 33554432 64x1048576:1 --variant=0 --variant=-1000 --variant=36 2>&1 | tail -30
ITERS=1 WARM=1 timeout 600 ncu --set full --clock-control none -k regex:k_scan -f -o /tmp/r02_scan python tools/scan_bench.py 67108864:None:float32 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02_scan.ncu-rep 2>/dev/null | grep -E "^==|time_duration|dram__bytes|warps_active|long_scoreboard_per|barrier_per|short_score|issue_active.avg|wait_per|no_instr|branch" | head -14
