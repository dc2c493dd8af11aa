mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_l0.py -m gpu -q -x -k "fps_bit_exact and pruned and (n3000 or model_size or ties_lattice_600 or n2049 or all_identical)" > /tmp/race.txt 2>&1
echo "racecheck lines naming our kernels: $(grep -c 'fps_pruned\|fps_register\|fps_cluster' /tmp/race.txt)"; grep -E "RACECHECK SUMMARY|passed|failed" /tmp/race.txt | head -3
grep -E "fps_pruned" /tmp/race.txt | head -5
timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -s -k "fused_wgrad or captured or golden" 2>&1 | grep -E "fused weight|passed|failed|Error|assert" | cut -c1-250
for f in 1 0; do
  CAMLI_FUSE_WGRAD=$f timeout 600 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_fuse$f.json 2> gpurun_out/bench_c5_fuse$f.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_c5_fuse$f.json')); print('fuse=$f c5 %.2f pairs/s %.1f ms loss %.4f' % (b['value'], b['ms_per_step'], b['config']['final_loss']))"
done
