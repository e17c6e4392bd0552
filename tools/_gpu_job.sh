mkdir -p gpurun_out
python -m pytest tests/test_gpu_assign.py tests/test_gpu_golden.py tests/test_zz_fullsize.py -m gpu -q -s 2>&1 | grep -E "ulp from the true|passed|failed|FAILED|Error|assert" | tail -30 > gpurun_out/r02_ulp_tests.log
cat gpurun_out/r02_ulp_tests.log
REPS=3 python tools/prof_cases.py cfg4
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln)
        for k in ('cfg4_transpose_view_f64','cfg3_sum_axis0','cumsum_flat_f32_2^26'): print(k, d.get(k))
"
python tools/ulp_report.py > gpurun_out/ulp_report_r02.json 2>/dev/null
ROWS=262144 REPS=1 timeout 900 ncu --set full --clock-control none -f -o /tmp/r02_cfg5 python tools/prof_cases.py cfg5map cfg5sum cfg5var cfg4 > gpurun_out/r02_ncu_cfg5.log 2>&1
python tools/ncu_summary.py /tmp/r02_cfg5.ncu-rep gpurun_out/r02_ncu_summary_cfg5_cfg4.csv > gpurun_out/r02_ncu_summary_cfg5_cfg4.txt 2>&1
ls -la /tmp/r02_cfg5.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_reduce.py tests/test_arg_norm.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_sanitizer_memcheck.log; echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py tests/test_gpu_reduce.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_sanitizer_racecheck.log; echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_assign.py -m gpu -x -q -k "tma or cfg4" 2>&1 | tail -8 > gpurun_out/r02_sanitizer_memcheck_tma.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log; tail -4 gpurun_out/r02_sanitizer_racecheck.log; tail -4 gpurun_out/r02_sanitizer_memcheck_tma.log; du -sh gpurun_out
