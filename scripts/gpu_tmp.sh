mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv_gemm or gru or motion or linear" 2>&1 | tail -2 | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -k "dense" 2>&1 | tail -2 | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_train_golden.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
timeout 300 python scripts/cg_time.py --header 2>&1 | tail -3
run() {  # name, args
  name=$1; shift
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); r=b['roofline']; print('$name value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency %.2f' % b['latency']['ms_per_pair'], 'conv fp32eq %.1f largest %.1f us' % (r['fp32_equivalent_TFLOPs'], r['largest']['avg_us']))"
}
run default
run c3 --workload c3 --pairs-per-step 4
