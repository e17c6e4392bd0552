mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu_f.log
tail -3 gpurun_out/r02_pytest_gpu_f.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_f.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'])
for k,v in d['other_configs'].items():
    if 'GBs' in v: print(k, v['GBs'], v['frac_of_measured_peak'], v['ms'], v['kernel'][:50])
P
