mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -s 2>&1 | grep -E "dense|passed|failed|Error|error|loss|worst|assert" | cut -c1-230 | tail -30
for r in library tcgen05; do
  CAMLI_TRAIN_DENSE=$r timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_$r.json 2> gpurun_out/bench_c5_$r.err
  tail -2 gpurun_out/bench_c5_$r.err
  python - <<PY
import json
b=json.load(open("gpurun_out/bench_c5_$r.json"))
print("$r", "c5 value %.2f pairs/s, ms/step %.1f" % (b["value"], b["ms_per_step"]), {k:b.get(k) for k in ("dtype",)}, b["config"].get("workload"))
PY
done
