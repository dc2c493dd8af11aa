mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/tmp_pytest.log
tail -4 gpurun_out/tmp_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err; tail -3 gpurun_out/tmp_bench.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/tmp_bench.json").read().strip().splitlines()[-1])
print('c2 value %.1f e2e %.1f latency %.2f ms' % (b['value'], b['e2e']['value'], b['latency']['ms_per_pair']))
PY
