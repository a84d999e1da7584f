#!/bin/bash
# round-end profiles: launch list of a bench step (shares), full captures of the three hot kernels
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 16384 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for k in nw_fast_kernel rank_kernel traceback_fast_tasks_kernel; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${k}_r01 python tools/stage_times.py 16384 --short 2>&1 | tail -1
done
ls -la gpurun_out/
