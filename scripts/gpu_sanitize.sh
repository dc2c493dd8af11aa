#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels added in the last session: pruned FPS, convex up-sampling
# forward / backward, PointConv grouping backward, strided selective-kernel tail.  Output: gpurun_out/sanitize.txt
mkdir -p gpurun_out
: > gpurun_out/sanitize.txt
for tool in memcheck racecheck; do
  for sel in "tests/test_gpu_l0.py -k (fps_bit_exact and pruned and (n3000 or model_size or ties_lattice_600 or n2049 or all_identical))" \
             "tests/test_gpu_grad.py -k (convex_upsample or pointconv_group_backward)" \
             "tests/test_gpu_ops.py -k sk_fusion_tail"; do
    echo "== $tool: $sel" >> gpurun_out/sanitize.txt
    # shellcheck disable=SC2086
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $(echo "$sel" | cut -d' ' -f1) -m gpu -q -x -k "$(echo "$sel" | cut -d' ' -f3-)" 2>&1 \
      | grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Invalid|hazard" | head -8 >> gpurun_out/sanitize.txt
  done
done
cat gpurun_out/sanitize.txt
