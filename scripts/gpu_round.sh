#!/bin/bash
# One GPU-box round: parity tests, bench, launch list, ncu captures. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -3 gpurun_out/kernels.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'corr2d_lookup_kernel|dw_gather_max_kernel|knn_warp_kernel|allpairs_tf32x3_kernel|corr3d_lookup_kernel' -c 12 -f -o gpurun_out/prof_kernels python scripts/profile_kernels.py --iters 1 > gpurun_out/ncu_kernels.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
