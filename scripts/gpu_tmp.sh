for keep in 0 1; do
echo "== CAMLI_LOOKUP_KEEP_L2=$keep"
CAMLI_LOOKUP_KEEP_L2=$keep timeout 600 python scripts/trace_forward.py --out gpurun_out/trace_keep$keep 2>&1 | grep -E "^graph|corr2d_lookup|span" | cut -c1-200
python - <<PY
import json
ev=json.load(open("gpurun_out/trace_keep${keep}_graph.json"))["traceEvents"]
ks=sorted([e for e in ev if e.get("cat")=="kernel" and "corr2d_lookup" in e["name"]], key=lambda e:e["ts"])
print("lookup durations per iteration:", [round(e["dur"],1) for e in ks])
PY
rm -f gpurun_out/trace_keep${keep}_graph.json
done
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "corr2d or lookup" 2>&1 | tail -2
