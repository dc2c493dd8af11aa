"""Kernel totals of one C5 training step (torch.profiler): where the backward spends its time.  Diagnosis only."""
import collections, json, os, sys
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camliflow_b200 import trainer  # noqa: E402
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402

H, W, N, iters, B = bench.WORKLOADS["c5"]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=iters)), seed=0).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
inp = bench.synthetic_inputs(B, H, W, N, 0)
g = torch.Generator().manual_seed(1)
inp["flow_2d"] = torch.randn(B, 2, H, W, generator=g) * 5
inp["flow_3d"] = torch.randn(B, 3, N, generator=g) * 0.1
inp = {k: v.to(dev) for k, v in inp.items()}
amp = torch.bfloat16 if os.environ.get("TRACE_AMP", "1") == "1" else None
for _ in range(2):
    trainer.train_step(model, opt, inp, autocast_dtype=amp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=False) as prof:
    trainer.train_step(model, opt, inp, autocast_dtype=amp)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:100]][0] += 1
        agg[e.name[:100]][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("total kernel time %.1f ms over %d kernels" % (tot / 1e3, sum(v[0] for v in agg.values())))
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6d %9.1f us %5.1f%%  %s" % (c, u, 100 * u / tot, n))
# CPU-side op totals (which autograd nodes dominate)
if os.environ.get("TRACE_OPS"):
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60)[:6000])
