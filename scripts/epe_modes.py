"""C2 end-point error against the reference golden under the precision modes the bench can run in
(diagnosis: which mode keeps EPE2D <= 1e-3 / EPE3D <= 1e-4).  Prints one JSON line per mode."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200 import ops  # noqa: E402
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402
from oracle import camliraft_oracle as co  # noqa: E402


def epe(a, b):
    return float(np.sqrt(((a - b) ** 2).sum(0)).mean())


def main():
    G = np.load(os.path.join(ROOT, "tests", "golden", "model_camliraft.npz"))
    inputs = {k: v.cuda() for k, v in co.synthetic_inputs(1, 540, 960, 8192, seed=0).items()}
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=12)), seed=0).cuda().eval()
    for allpairs, tf32, cl in [("tcgen05", False, True), ("cublas", False, True), ("tcgen05", True, True)]:
        ops.ALLPAIRS_IMPL = allpairs
        ops.SINGLE_PASS_INFERENCE = tf32          # one tf32 product per element in the dense kernel instead of three
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        m = model.to(memory_format=torch.channels_last if cl else torch.contiguous_format)
        m.channels_last = cl
        with torch.no_grad():
            out = m(inputs)
        d2 = np.sqrt(((out["flow_2d"][0, :, ::8, ::8].cpu().numpy() - G["c2_kernel_flow2d"]) ** 2).sum(0)).ravel()
        d3 = np.sqrt(((out["flow_3d"][0, :, ::4].cpu().numpy() - G["c2_kernel_flow3d"]) ** 2).sum(0)).ravel()
        dist = lambda d: {"mean": float(d.mean()), "median": float(np.median(d)), "p99": float(np.percentile(d, 99)),   # noqa: E731
                          "max": float(d.max()), "share_of_sum_in_top_1pct": float(np.sort(d)[-max(1, d.size // 100):].sum() / d.sum())}
        print(json.dumps({"allpairs": allpairs, "dense_single_pass_tf32": tf32, "channels_last": cl, "epe2d": dist(d2), "epe3d": dist(d3)}))


if __name__ == "__main__":
    main()
