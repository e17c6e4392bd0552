mkdir -p gpurun_out
python -m pytest tests/test_gpu_reduce.py tests/test_nan_functions.py tests/test_arg_norm.py tests/test_zz_fullsize.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_pdl_pytest.log
python tools/split_sweep.py > gpurun_out/r02_split_sweep3.log 2>&1
tail -3 gpurun_out/r02_pdl_pytest.log
cat gpurun_out/r02_split_sweep3.log | cut -c1-120
