#!/bin/bash
# Final round evidence (cheap): bench line + quick ncu of the kernels changed since the r1_v3 captures.
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
kill $SMI
tail -c 400 gpurun_out/bench.json
timeout 600 ncu --set full --clock-control none -k regex:'knn_warp_kernel|backwarp3d_kernel|dw_gather_max|corr2d_lookup_kernel' -c 12 -f -o /tmp/prof_final python scripts/profile_kernels.py --iters 1 > gpurun_out/ncu_kernels.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
rm -f gpurun_out/launches.csv gpurun_out/prof_kernels_details.csv gpurun_out/prof_source_*.csv
ls -la gpurun_out | head -20
