mkdir -p gpurun_out
python -m pytest tests/test_gpu_reduce.py tests/test_nan_functions.py tests/test_arg_norm.py tests/test_zz_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_decomp_pytest.log
tail -5 gpurun_out/r02_decomp_pytest.log
python tools/reduce_bench.py 2>&1 | tee gpurun_out/r02_reduce_bench3.log | cut -c1-150
