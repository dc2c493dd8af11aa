#!/bin/bash
# Final evidence run of the round: gpu_profile.sh (bench line + ncu launch list + ncu --set full of the hot kernels), then the
# bench lines of the other workloads, the training A/B (hand-written vs library dense layers) and the L0 microbench against the
# reference's own kernels.  Everything lands in gpurun_out/ as text; scripts/summarise_profiles.py turns it into profiles/.
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-200 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200
bash scripts/gpu_profile.sh
for w in c3 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --pairs-per-step 4 --no-cpu-baseline --no-training-block > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err | cut -c1-200
done
timeout 600 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 600 python bench.py --workload c5 --train-dense library --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_library.json 2> gpurun_out/bench_c5_library.err
timeout 600 python bench.py --workload c5 --train-precision fp32 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_fp32.json 2> gpurun_out/bench_c5_fp32.err
timeout 600 python bench.py --workload c5 --train-precision fp32 --train-dense library --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_fp32_library.json 2> gpurun_out/bench_c5_fp32_library.err
timeout 300 python scripts/microbench_l0.py > gpurun_out/microbench_l0.log 2>&1; tail -3 gpurun_out/microbench_l0.log | cut -c1-200
timeout 600 python scripts/trace_train.py 2>&1 | grep -v Warning | tail -48 | cut -c1-150 > gpurun_out/trace_train.txt
timeout 300 python scripts/cg_time.py --header > gpurun_out/cg_time.txt 2>&1; cat gpurun_out/cg_time.txt
CG_TIME_SMALL=1 timeout 300 python scripts/cg_time.py --header > gpurun_out/cg_time_small.txt 2>&1
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1
python scripts/trace_iteration.py > gpurun_out/trace_iteration.txt 2>&1; head -9 gpurun_out/trace_iteration.txt
python - <<'PY'
import json
for f in ("bench", "bench_c3", "bench_c4", "bench_c5", "bench_c5_library", "bench_c5_fp32", "bench_c5_fp32_library"):
    try:
        b = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        extra = ""
        if "latency" in b: extra = " e2e %.1f latency %.2f ms" % (b["e2e"]["value"], b["latency"]["ms_per_pair"])
        if b.get("roofline"): extra += " | %s frac %.3f" % (b["roofline"]["kernel"], b["roofline"]["frac"])
        if b.get("training"): extra += " | training %.2f pairs/s" % b["training"]["value"]
        if b.get("cpu_baseline"): extra += " | cpu %.3f" % b["cpu_baseline"]["value"]
        print("%-24s value %.2f %s%s" % (f, b["value"], b["unit"], extra))
    except Exception as e:
        print(f, "ERR", e)
PY
ls gpurun_out | wc -l; du -sh gpurun_out
