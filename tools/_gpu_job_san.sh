mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py tests/test_arg_norm.py tests/test_gpu_runtime.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02b_sanitizer_memcheck.log; echo "memcheck rc=$?" >> gpurun_out/r02b_sanitizer_memcheck.log
tail -6 gpurun_out/r02b_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x -k "decomposed or unaligned or shapes_bit_exact" 2>&1 | tail -8 > gpurun_out/r02b_sanitizer_racecheck.log; echo "racecheck rc=$?" >> gpurun_out/r02b_sanitizer_racecheck.log
tail -6 gpurun_out/r02b_sanitizer_racecheck.log
