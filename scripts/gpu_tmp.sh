mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "sk_fusion" 2>&1 | tail -4
timeout 200 python scripts/sk_time.py 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_sk.json 2> gpurun_out/bench_sk.err
tail -2 gpurun_out/bench_sk.err
python - <<PY
import json
b=json.load(open("gpurun_out/bench_sk.json")); r=b["roofline"]
print("value %.1f e2e %.1f latency %.2f ms sync %.1f" % (b["value"], b["e2e"]["value"], b["latency"]["ms_per_pair"], b["e2e"]["synchronous"]["value"]), r["kernel"], "frac %.3f fp32eq %.1f" % (r["frac"], r["fp32_equivalent_TFLOPs"]), r.get("timing"))
PY
