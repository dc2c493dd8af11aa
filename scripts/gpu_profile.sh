#!/bin/bash
# Evidence run for profiles/: bench line, ncu launch list of the same command, ncu --set full of the hot kernels.
# Everything that travels back is text (the .ncu-rep is converted on the box: gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
kill $SMI
tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
CAMLI_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-training-block --concurrent 1 > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'corr2d_lookup_kernel|dw_gather_max|knn_warp_kernel|allpairs_tf32x3_kernel|corr3d_lookup_kernel|conv_gemm_tf32x3_kernel|conv_wgrad_tf32x3_kernel|transpose_split_kernel|fps_cluster_async|fps_pruned|convex_upsample|pointconv_group|sk_|clfm_interp' \
    -c 70 -f -o /tmp/prof_kernels python scripts/profile_kernels.py --iters 1 > gpurun_out/ncu_kernels.log 2>&1
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
ncu -i /tmp/prof_kernels.ncu-rep --page details --csv > gpurun_out/prof_kernels_details.csv 2>/dev/null
for k in corr2d_lookup_kernel dw_gather_max_kernel conv_gemm_tf32x3_kernel; do
  ncu -i /tmp/prof_kernels.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > gpurun_out/prof_source_$k.csv 2>/dev/null
done
ls -la /tmp/prof_kernels.ncu-rep gpurun_out/ | tail -20; du -sh gpurun_out
