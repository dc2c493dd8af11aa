mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -k "convex" 2>&1 | tail -5 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_train_golden.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -k "golden or captured or training_step" 2>&1 | tail -3 | cut -c1-200
run() {  # name, args
  name=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); print('$name value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency', (b.get('latency') or {}).get('ms_per_pair'), 'training', (b.get('training') or {}).get('value'))"
}
run default
run c3 --workload c3 --pairs-per-step 4 --no-training-block
run c5 --workload c5 --steps 6
