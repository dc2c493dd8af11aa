#!/bin/bash
# Round-2 check: full GPU test-suite, then the bench lines of every workload (short runs).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench_c2.json"))
    print("c2 value %.1f e2e %.1f latency %.2f ms sync %.1f" % (b["value"], b["e2e"]["value"], b["latency"]["ms_per_pair"], b["e2e"]["synchronous"]["value"]))
    r=b["roofline"]; print("roofline", r["kernel"], r["bound"], "%.1f/%.1f frac %.3f fp32eq %.1f share %.2f" % (r["achieved"], r["peak"], r["frac"], r.get("fp32_equivalent_TFLOPs",0), r["share_of_step"]))
    print("corr_lookup", r.get("corr_lookup")); print("knn_gather", r.get("knn_gather"))
    print("training", {k:b["training"][k] for k in ("value","ms_per_step","dtype")}, b["training"]["config"]["final_loss"])
    print("cpu", b.get("cpu_baseline"))
except Exception as e: print("ERR", e)
PY
for w in c3 c4; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --pairs-per-step 4 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -3 gpurun_out/bench_$w.err
python - <<PY
import json
try:
    b=json.load(open("gpurun_out/bench_$w.json"))
    print("$w value %.1f e2e %.1f latency %.2f ms/pair" % (b["value"], b["e2e"]["value"], b["latency"]["ms_per_pair"]), b["roofline"]["kernel"], b["roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
done
timeout 300 python scripts/conv_gemm_timeline.py 2>&1 | grep "==\|setup done" > gpurun_out/cg_timeline.log; cat gpurun_out/cg_timeline.log
timeout 600 python scripts/diag_c4.py 1 2>&1 | grep -v Warn > gpurun_out/c4_divergence.txt; head -5 gpurun_out/c4_divergence.txt
