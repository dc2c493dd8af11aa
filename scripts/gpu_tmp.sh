mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/tmp_pytest.log
tail -4 gpurun_out/tmp_pytest.log; grep "camlipwc" gpurun_out/tmp_pytest.log | head
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --pairs-per-step 4 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_c3.json"))
print("c3 value %.1f e2e %.1f latency %.2f ms/pair" % (b["value"], b["e2e"]["value"], b["latency"]["ms_per_pair"]), b["roofline"]["kernel"], b["roofline"]["frac"])
PY
