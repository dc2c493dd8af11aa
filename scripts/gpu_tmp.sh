mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/tmp_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-training-block > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
tail -5 gpurun_out/tmp_pytest.log; tail -5 gpurun_out/tmp_bench.err
