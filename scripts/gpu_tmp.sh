mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -s -k "single_pass or training or captured or stride2" 2>&1 | grep -E "single|passed|failed|Error|error|assert|worst" | cut -c1-220 | tail -12
timeout 600 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python -c "
import json; b=json.load(open('gpurun_out/bench_c5.json')); print('bf16 autocast, bf16 wgrad: c5 %.2f pairs/s %.1f ms loss %.4f' % (b['value'], b['ms_per_step'], b['config']['final_loss']))"
timeout 600 python scripts/trace_train.py 2>&1 | grep -v Warning | grep -E "total kernel|transpose_split|conv_wgrad|conv_gemm" | cut -c1-150
