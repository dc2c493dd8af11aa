"""Per-op CUDA-event timings of the L0 kernels against the reference's own kernels
(oracle/_ref) on the same box.  Writes gpurun_out/microbench_l0.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _util  # noqa: E402
from camliflow_b200 import csrc, native  # noqa: E402
from camliflow_b200.csrc import wrapper  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    have_ref = _util.ref_lib() is not None
    res = {}

    def row(name, ours, ref=None):
        res[name] = {"ours_us": round(timeit(ours), 2)}
        if have_ref and ref is not None:
            res[name]["ref_us"] = round(timeit(ref, iters=5, warmup=1), 2)
        print(name, res[name], flush=True)

    for B in (2, 8):
        pc = _util.synthetic_pc(B, 8192, seed=0).to(dev)
        row("fps_%dx8192_s4096" % B, lambda: csrc.furthest_point_sampling(pc, 4096), lambda: _util.ref_fps(pc, 4096))
        for mode, tag in ((2, "cluster_async"), (0, "single_cta_unpruned")):      # the other kernels behind the same entry point
            old = native.lib().camli_fps_set_cluster_path(mode)
            try:
                row("fps_%dx8192_s4096_%s" % (B, tag), lambda: csrc.furthest_point_sampling(pc, 4096))
            finally:
                native.lib().camli_fps_set_cluster_path(old)
    pc = _util.synthetic_pc(2, 4096, seed=0).to(dev)
    row("fps_2x4096_s2048", lambda: csrc.furthest_point_sampling(pc, 2048), lambda: _util.ref_fps(pc, 2048))
    for (n, m, k, D) in [(2048, 2048, 32, 3), (2048, 2048, 16, 3), (4096, 8192, 16, 3), (2048, 4096, 16, 3),
                         (8192, 2048, 3, 3), (1024, 2048, 3, 3), (2048, 256, 16, 3), (8160, 2048, 1, 2),
                         (34560, 4096, 1, 2)]:
        inp = _util.rand_cloud(1, m, D, seed=1).to(dev)
        qry = _util.rand_cloud(1, n, D, seed=2).to(dev)
        row("knn%dd_n%d_m%d_k%d" % (D, n, m, k), lambda: wrapper._k_nearest_neighbor_cuda(inp, qry, k),
            lambda: _util.ref_knn(inp, qry, k))
    for (B, C, H, W) in [(4, 32, 144, 240), (4, 64, 72, 120), (4, 96, 36, 60), (4, 128, 18, 30), (4, 192, 9, 15)]:
        a = torch.rand((B, H, W, C), device=dev)
        b = torch.rand((B, H, W, C), device=dev)
        go = torch.rand((B, 81, H, W), device=dev)
        row("corr_fwd_%dx%dx%dx%d" % (B, C, H, W), lambda: wrapper._correlation_forward_cuda(a, b, 4),
            lambda: _util.ref_corr_fwd(a, b, 4))
        row("corr_bwd_%dx%dx%dx%d" % (B, C, H, W), lambda: wrapper._correlation_backward_cuda(go, a, b, 4),
            lambda: _util.ref_corr_bwd(go, a, b, 4))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench_l0.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
