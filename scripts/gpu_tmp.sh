mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
timeout 600 python scripts/epe_modes.py > gpurun_out/epe_modes.jsonl 2> gpurun_out/epe_modes.err; python - <<'PY'
import json
for l in open("gpurun_out/epe_modes.jsonl"):
    d=json.loads(l); print(d["allpairs"], "single_pass_tf32" if d["dense_single_pass_tf32"] else "3xTF32", "EPE2D %.3e EPE3D %.3e" % (d["epe2d"]["mean"], d["epe3d"]["mean"]))
PY
timeout 600 python bench.py --conv-precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; tail -2 gpurun_out/bench_tf32.err | cut -c1-200
python -c "
import json; b=json.load(open('gpurun_out/bench_tf32.json')); print('tf32 inference: value %.1f e2e %.1f latency %.2f ms' % (b['value'], b['e2e']['value'], b['latency']['ms_per_pair']), b['dtype'])"
timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python -c "
import json; b=json.load(open('gpurun_out/bench_c5.json')); print('bf16 tcgen05 (padded layers) c5 %.2f pairs/s %.1f ms' % (b['value'], b['ms_per_step']))"
