#!/bin/bash
# GPU box round: parity tests, EPE per precision mode, kernel microbench, bench, kernel timeline.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|EPE|Error" gpurun_out/pytest_gpu.log
timeout 600 python scripts/epe_modes.py > gpurun_out/epe_modes.jsonl 2> gpurun_out/epe_modes.err; cat gpurun_out/epe_modes.jsonl; tail -3 gpurun_out/epe_modes.err
timeout 300 python scripts/profile_kernels.py --iters 3 > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -3 gpurun_out/kernels.err
python - <<'PY'
import json
k=json.load(open("gpurun_out/kernels.json"))
for n,v in k.items():
    if n.endswith("/min"): print("%-48s %s"%(n[:-4], v))
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("fp32",):
    try:
        b=json.load(open("gpurun_out/bench_%s.json"%f)); print(f, b["value"], b["ms_per_step"], b["e2e"]["value"], b["gpu_launches"], b.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1; tail -45 gpurun_out/trace.log | cut -c1-150
