mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_l0.py -m gpu -q -x -k fps 2>&1 | tail -3 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-200
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); print('$name value %.1f e2e %.1f latency %.2f ms sync %.1f' % (b['value'], b['e2e']['value'], b['latency']['ms_per_pair'], b['e2e']['synchronous']['value']))"
}
run default X=1
run wave100 CAMLI_LATENCY_WAVE=100
run wave74 CAMLI_LATENCY_WAVE=74
run fps2 CAMLI_FPS_PATH=2
run noaux CAMLI_AUX_STREAMS=0
timeout 600 python scripts/trace_forward.py > gpurun_out/trace.log 2>&1
python scripts/trace_iteration.py > gpurun_out/trace_iteration.txt 2>&1; head -12 gpurun_out/trace_iteration.txt
