mkdir -p gpurun_out
python -c "import torch; print('priority range', torch.cuda.Stream.priority_range())"
run() {  # name, env
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_$name.json')); print('$name value %.2f e2e %.2f' % (b['value'], b['e2e']['value']), 'latency %.3f' % b['latency']['ms_per_pair'], 'sync %.1f' % b['e2e']['synchronous']['value'])"
}
run prio_m1_s2 CAMLI_MAIN_PRIORITY=-1 CAMLI_SIDE_PRIORITY=-2
run default X=1
