#!/bin/bash
# ncu launch list of the timed single-engine region of the bench command (eager, no graph): per-launch device times.
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep gpurun_out/prof_kernels_raw.csv
CAMLI_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --concurrent 1 > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
