"""Runs the REFERENCE CamLiRAFT model (imported from /root/reference; build container only) on
seeded inputs with name-seeded weights and writes (sub-sampled) outputs to
tests/golden/model_camliraft.npz.  Two index semantics per case:

  *_fallback : the reference exactly as it runs on CPU tensors (pure-torch FPS / k-NN fallbacks);
  *_kernel   : the reference's Python graph with its FPS / k-NN calls answered by
               oracle/kernels_oracle.c, the restatement of the reference's CUDA kernels that
               tests/golden/l0_reference_cuda.npz pins bit-exact against those kernels run on a
               B200 -- i.e. what the reference computes on a GPU.

    python tests/golden/make_golden_model.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
from oracle import camliraft_oracle as co  # noqa: E402

# name: (H, W, N, n_iters, seed, stride2d, stride3d)
CASES = {
    "small": (160, 224, 8192, 3, 11, 4, 4),
    "c2": (540, 960, 8192, 12, 0, 8, 4),
}


def patch_indices(models, on):
    """Answer the reference's FPS / k-NN calls with the kernel-semantics C oracle."""
    import importlib
    mods = [importlib.import_module("models." + m) for m in
            ("utils", "point_conv", "clfm", "camliraft_core", "camliraft_l_core", "camlipwc_core", "camlipwc_l_core")]
    wrapper = importlib.import_module("models.csrc.wrapper")
    if not on:
        for m in mods:
            if hasattr(m, "k_nearest_neighbor"):
                m.k_nearest_neighbor = wrapper.k_nearest_neighbor
        mods[0].furthest_point_sampling = wrapper.furthest_point_sampling
        return

    def knn(input_xyz, query_xyz, k, cpp_impl=True):
        if input_xyz.shape[1] > 3:
            input_xyz, query_xyz = input_xyz.transpose(1, 2), query_xyz.transpose(1, 2)
        return co.knn(input_xyz.detach(), query_xyz.detach(), k, "kernel")    # (indices carry no gradient)

    def fps(xyz, n_samples, cpp_impl=True):
        return co.fps(xyz.detach().contiguous(), n_samples, "kernel")

    for m in mods:
        if hasattr(m, "k_nearest_neighbor"):
            m.k_nearest_neighbor = knn
    mods[0].furthest_point_sampling = fps


def main():
    torch.set_num_threads(8)
    models = rh.load_reference()
    P = co.make_params(co.param_spec("camliraft"), seed=0)
    out = {}
    for name, (H, W, N, iters, seed, s2, s3) in CASES.items():
        net = models.camliraft.CamLiRAFT(rh.camliraft_cfg(n_iters=iters)).eval()
        missing = net.load_state_dict(P, strict=True)
        print(name, missing)
        inputs = co.synthetic_inputs(1, H, W, N, seed)
        for mode in ("fallback", "kernel"):
            patch_indices(models, mode == "kernel")
            with torch.no_grad():
                res = net(inputs)
            f2, f3 = res["flow_2d"], res["flow_3d"]
            print(name, mode, tuple(f2.shape), tuple(f3.shape), float(f2.abs().mean()), float(f3.abs().mean()),
                  float(f2.abs().max()), float(f3.abs().max()))
            out["%s_%s_flow2d" % (name, mode)] = f2[0, :, ::s2, ::s2].numpy().astype(np.float32)
            out["%s_%s_flow3d" % (name, mode)] = f3[0, :, ::s3].numpy().astype(np.float32)
        patch_indices(models, False)
    np.savez_compressed(os.path.join(HERE, "model_camliraft.npz"), **out)

    # ---- CamLiPWC (BASELINE config[2] shape; kernel index semantics only)
    P = co.make_params(co.param_spec("camlipwc"), seed=0)
    net = models.camlipwc.CamLiPWC(rh.camlipwc_cfg()).eval()
    print("camlipwc", net.load_state_dict(P, strict=True))
    out = {}
    for name, (H, W, N, seed, s2, s3) in {"small": (128, 192, 8192, 21, 4, 4), "c3": (540, 960, 8192, 1, 8, 4)}.items():
        inputs = co.synthetic_inputs(1, H, W, N, seed)
        patch_indices(models, True)
        with torch.no_grad():
            res = net(inputs)
        patch_indices(models, False)
        f2, f3 = res["flow_2d"], res["flow_3d"]
        print("camlipwc", name, tuple(f2.shape), tuple(f3.shape), float(f2.abs().mean()), float(f3.abs().mean()))
        out["%s_kernel_flow2d" % name] = f2[0, :, ::s2, ::s2].numpy().astype(np.float32)
        out["%s_kernel_flow3d" % name] = f3[0, :, ::s3].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "model_camlipwc.npz"), **out)


if __name__ == "__main__":
    main()
