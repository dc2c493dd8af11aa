mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -k "dense or training or captured" 2>&1 | tail -3
timeout 600 python scripts/trace_train.py 2>&1 | grep -v Warning | grep -E "total kernel|transpose_split|conv_wgrad|conv_gemm" | cut -c1-150
timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_tcgen05.json 2> gpurun_out/bench_c5_tcgen05.err
python -c "
import json; b=json.load(open('gpurun_out/bench_c5_tcgen05.json')); print('bf16 tcgen05 c5 %.2f pairs/s %.1f ms' % (b['value'], b['ms_per_step']))"
