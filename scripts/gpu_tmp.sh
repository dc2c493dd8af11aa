mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
run() {  # name, env
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); print('$name value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency %.3f' % b['latency']['ms_per_pair'], 'sync %.1f' % b['e2e']['synchronous']['value'])"
}
run default X=1
run default2 X=1
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1
python scripts/trace_iteration.py > gpurun_out/trace_iteration.txt 2>&1; head -9 gpurun_out/trace_iteration.txt
