mkdir -p gpurun_out /tmp/ncu
ROWS=262144 REPS=1 ncu --set full --clock-control none -o /tmp/ncu/r02b_cfg5 -f python tools/prof_cases.py cfg5sum cfg5var cfg5map > gpurun_out/r02b_ncu_cfg5.log 2>&1
REPS=1 ncu --set full --clock-control none -o /tmp/ncu/r02b_cfg3 -f python tools/prof_cases.py cfg3a0 cfg3max0 > gpurun_out/r02b_ncu_cfg3.log 2>&1
for c in cfg5 cfg3; do
  python tools/ncu_summary.py /tmp/ncu/r02b_$c.ncu-rep > gpurun_out/r02b_ncu_summary_$c.txt
  ncu -i /tmp/ncu/r02b_$c.ncu-rep --page raw --csv | gzip > gpurun_out/r02b_ncu_raw_$c.csv.gz
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r02b_bench_under_ncu.log 2>&1
ls -la gpurun_out/ /tmp/ncu
