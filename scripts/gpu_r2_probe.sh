#!/bin/bash
# Round-2 opening probe: per-shape launch timings, L0 micro-benchmark against the reference kernels, stream timeline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -2
timeout 600 python scripts/trace_conv_shapes.py > gpurun_out/conv_shapes.log 2>&1; head -80 gpurun_out/conv_shapes.log
timeout 600 python scripts/microbench_l0.py > gpurun_out/microbench_l0.log 2>&1; tail -40 gpurun_out/microbench_l0.log
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1; tail -60 gpurun_out/trace.log | cut -c1-160
