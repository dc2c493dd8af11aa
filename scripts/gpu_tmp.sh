mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_mask.json 2> gpurun_out/bench_mask.err
python -c "
import json; b=json.load(open('gpurun_out/bench_mask.json')); print('value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency %.3f' % b['latency']['ms_per_pair'], 'sync %.1f' % b['e2e']['synchronous']['value'])"
