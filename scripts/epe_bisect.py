"""Which component sets the C2 end-point error?  Runs the C2 parity case with parts of the tensor-core path
switched back to cuDNN (diagnosis)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200 import tc, raft_core  # noqa: E402
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402
from oracle import camliraft_oracle as co  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "model_camliraft.npz"))
inputs = {k: v.cuda() for k, v in co.synthetic_inputs(1, 540, 960, 8192, seed=0).items()}
model = seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=12)), seed=0).cuda().eval()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
orig_fused_inf = raft_core._fused_inference
orig_fused = tc.fused


def run(tag, cl):
    m = model.to(memory_format=torch.channels_last if cl else torch.contiguous_format)
    m.channels_last = cl
    with torch.no_grad():
        out = m(inputs)
    d2 = np.sqrt(((out["flow_2d"][0, :, ::8, ::8].cpu().numpy() - G["c2_kernel_flow2d"]) ** 2).sum(0))
    d3 = np.sqrt(((out["flow_3d"][0, :, ::4].cpu().numpy() - G["c2_kernel_flow3d"]) ** 2).sum(0))
    print(json.dumps({"variant": tag, "channels_last": cl, "epe2d": float(d2.mean()), "epe3d": float(d3.mean()),
                      "max2d": float(d2.max())}), flush=True)


run("all tensor-core layers on", True)
raft_core._fused_inference = lambda x: False
run("encoder through cuDNN modules (conv, BN, ReLU), rest tensor-core", True)
run("encoder through cuDNN modules (conv, BN, ReLU), rest tensor-core", False)
raft_core._fused_inference = orig_fused_inf
tc.ENABLED = False
run("no tensor-core layers (cuDNN/cuBLAS everywhere), fused cuDNN encoder epilogues", True)
run("no tensor-core layers (cuDNN/cuBLAS everywhere), fused cuDNN encoder epilogues", False)
raft_core._fused_inference = lambda x: False
run("no tensor-core layers, encoder through modules", True)
run("no tensor-core layers, encoder through modules", False)
