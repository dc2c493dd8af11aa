mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
run() {  # name, env, args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); print('$name value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency', (b.get('latency') or {}).get('ms_per_pair'), 'sync', b['e2e']['synchronous']['value'])"
}
run default X=1
run prio0 CAMLI_MAIN_PRIORITY=0
run default2 X=1
run nopipe CAMLI_PIPELINE_3D=0
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1
python scripts/trace_iteration.py > gpurun_out/trace_iteration.txt 2>&1; head -9 gpurun_out/trace_iteration.txt
