mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --pairs-per-step 4 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err
python - <<'PY'
import json
for w in ("c2","c4"):
    b=json.load(open("gpurun_out/bench_%s.json"%w)); r=b["roofline"]
    print(w, "value %.1f e2e %.1f latency %.2f ms/pair" % (b["value"], b["e2e"]["value"], b["latency"]["ms_per_pair"]), r["kernel"], "frac %.3f fp32eq %.1f"%(r["frac"], r["fp32_equivalent_TFLOPs"]), "largest", {k:round(v,3) for k,v in r["largest"].items()})
    for k in ("corr_lookup","knn_gather"):
        if k in r: print("   ", k, {kk:(round(v,3) if isinstance(v,float) else v) for kk,v in r[k].items() if kk not in ("ncu","note","largest")}); print("       largest", {kk:round(v,3) for kk,v in r[k]["largest"].items()})
PY
