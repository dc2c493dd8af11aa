timeout 900 python -m pytest tests/test_gpu_grad.py -m gpu -q -x -s -k "captured or golden" 2>&1 | grep -v Warn | tail -8
timeout 900 python bench.py --workload c5 --steps 4 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print('c5 graph: %.2f pairs/s %.1f ms/step e2e %.2f loss %.3f launches %d' % (b['value'], b['ms_per_step'], b['e2e']['value'], b['config']['final_loss'], b['gpu_launches']))"
timeout 900 python bench.py --workload c5 --steps 4 --warmup 3 --no-graph 2>&1 | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print('c5 eager: %.2f pairs/s %.1f ms/step e2e %.2f loss %.3f' % (b['value'], b['ms_per_step'], b['e2e']['value'], b['config']['final_loss']))"
