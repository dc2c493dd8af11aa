"""Copies the per-workload bench lines of scripts/gpu_final.sh from gpurun_out/ into profiles/ (tracked evidence)."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2_final"
for name in ("bench_c3", "bench_c4", "bench_c5", "bench_c5_library", "bench_c5_fp32", "bench_c5_fp32_library"):
    src = os.path.join(OUT, name + ".json")
    if os.path.exists(src) and os.path.getsize(src):
        line = json.loads(open(src).read().strip().splitlines()[-1])
        json.dump(line, open(os.path.join(PROF, "%s_%s.json" % (tag, name)), "w"), indent=1)
        print(name, round(line["value"], 2), line["unit"])
for name in ("microbench_l0.json", "cg_time.txt", "cg_time_small.txt", "trace_train.txt", "trace_iteration.txt", "cg_timeline.log"):
    src = os.path.join(OUT, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "%s_%s" % (tag, name)))
