mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 --quick > gpurun_out/r02_n8x_$name.json 2> gpurun_out/r02_n8x_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r02_n8x_$name.json').read().strip().splitlines()[-1])
    print('$name', d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name', 'failed', e)
P
}
run default A=1
run nofork XTB_BENCH_NO_FORK=1
run split55 XTB_REDUCE_SPLIT=55
run nofork_split55 XTB_BENCH_NO_FORK=1 XTB_REDUCE_SPLIT=55
run nopdl XTB_NO_PDL=1
run default2 A=1
