"""Which lines of the host code run the library (ATen) ops that are left in a C2 forward: a TorchDispatchMode counts every aten
op that computes on the GPU by the innermost camliflow_b200 frame that issued it.  Diagnosis only."""
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

VIEW_OPS = ("aten.view", "aten.permute", "aten.transpose", "aten.slice", "aten.select", "aten.expand", "aten.detach", "aten.alias",
            "aten.unsqueeze", "aten.squeeze", "aten.t.", "aten.as_strided", "aten.split", "aten.unbind", "aten.reshape",
            "aten._unsafe_view", "aten.empty", "aten.sym_", "aten.unfold", "aten.chunk", "aten.narrow", "aten.lift_fresh",
            "aten.new_empty", "aten._reshape_alias")


class Count(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.agg = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = str(func)
        if not any(name.startswith(v) for v in VIEW_OPS):
            frame = "?"
            for f in reversed(traceback.extract_stack(limit=40)):
                if "camliflow_b200" in f.filename and "scripts" not in f.filename:
                    frame = "%s:%d %s" % (f.filename.split("camliflow_b200/")[-1], f.lineno, (f.line or "").strip()[:70])
                    break
            self.agg[(frame, name.replace("aten.", ""))] += 1
        return out


dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
model = bench.build_model("c2").to(dev).eval().to(memory_format=torch.channels_last)
model.channels_last = True
H, W, N, iters, B = bench.WORKLOADS["c2"]
inp = {k: v.to(dev) for k, v in bench.synthetic_inputs(B, H, W, N, 0).items()}
with torch.no_grad():
    model(inp)
    torch.cuda.synchronize()
    with Count() as c:
        model(inp)
    torch.cuda.synchronize()
print("aten compute ops in one forward: %d" % sum(c.agg.values()))
for (frame, op), n in c.agg.most_common(70):
    print("%4d  %-22s %s" % (n, op[:22], frame))
