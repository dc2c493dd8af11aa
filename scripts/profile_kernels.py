"""Runs the hot hand-written kernels at the C2 (960x540 + 8192 pts, B=1) sizes a few times, for
`ncu --set full -k regex:...` captures and quick CUDA-event timings (prints a JSON summary).

    python scripts/profile_kernels.py [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200 import native, ops  # noqa: E402
from camliflow_b200.csrc import furthest_point_sampling, k_nearest_neighbor  # noqa: E402
from camliflow_b200.mlp import MLP2d  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    H, W, N = 68, 120, 2048
    flush = torch.zeros(192 * 1024 * 1024 // 4, device=dev)

    f1 = torch.randn(1, 256, H, W, generator=g).to(dev)
    f2 = torch.randn(1, 256, H, W, generator=g).to(dev)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = (torch.stack([xs, ys], 0)[None] + torch.randn(1, 2, H, W, generator=g) * 2).to(dev)
    xyz = ((torch.rand(1, 3, N, generator=g) - 0.5) * 10).to(dev)
    pc = ((torch.rand(2, 8192, 3, generator=g) - 0.5) * 10).to(dev)
    feat = torch.randn(1, N, 128, generator=g).to(dev)
    wn = MLP2d(3, [8, 32, 128], act="relu").to(dev)
    flow = (torch.randn(1, 3, N, generator=g) * 0.1).to(dev)
    up_flow = torch.randn(1, 2, H, W, generator=g).to(dev)
    up_mask = torch.randn(1, 576, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)

    with torch.no_grad():
        pyr = ops.corr2d_build(f1, f2, 4)
        nbr = k_nearest_neighbor(xyz, xyz, 32)
        wc = ops.pointconv_dw_weights(xyz, xyz, nbr, 32, wn)
        cases = {
            "corr2d_build(allpairs+pool)": lambda: ops.corr2d_build(f1, f2, 4),
            "corr2d_lookup": lambda: ops.corr2d_lookup(pyr, coords, 4),
            "dw_gather_max_k32_O128": lambda: ops.pointconv_dw_gather_max(feat, wc, nbr, 32),
            "dw_gather_max_k16_O128": lambda: ops.pointconv_dw_gather_max(feat, wc[:, :, :16].contiguous(), nbr, 16),
            "dw_weights_k32_O128": lambda: ops.pointconv_dw_weights(xyz, xyz, nbr, 32, wn),
            "knn_2048x2048_k32": lambda: k_nearest_neighbor(xyz, xyz, 32),
            "knn_2048x2048_k16": lambda: k_nearest_neighbor(xyz, xyz, 16),
            "backwarp_3d_2048": lambda: ops.backwarp_3d(xyz, xyz, flow, 3),
            "fps_2x8192_s4096": lambda: furthest_point_sampling(pc, 4096),
            "convex_upsample_x8_68x120": lambda: ops.convex_upsample(up_flow, up_mask, 8, 0.25),
        }
        # tensor-core linear / convolution kernel at the shapes of the model (B, H, W, Cin, Cout, kh, kw)
        for (cb, ch, cw, ci, co, kh, kw) in [(1, 1, 2048, 384, 128, 1, 1), (1, 1, 2048, 128, 128, 1, 1),
                                             (1, 1, 4096, 1584, 96, 1, 1), (1, 68, 120, 324, 256, 1, 1),
                                             (1, 68, 120, 256, 192, 3, 3), (1, 68, 120, 384, 256, 1, 5),
                                             (1, 68, 120, 128, 256, 3, 3), (2, 136, 240, 64, 64, 3, 3),
                                             (2, 136, 240, 256, 64, 1, 1), (2, 136, 240, 64, 256, 1, 1),
                                             (2, 68, 120, 128, 128, 3, 3), (2, 68, 120, 512, 128, 1, 1)]:
            xin = torch.randn(cb, ch, cw, ci, generator=g).to(dev)
            wt = (torch.randn(co, kh * kw * ci, generator=g) / (kh * kw * ci) ** 0.5).to(dev)
            w_hi, w_lo, _ = ops.tc_weight([wt], lambda wt=wt: (wt, None))
            cases["conv_gemm_%dx%dx%d_%d->%d_%dx%d" % (cb, ch, cw, ci, co, kh, kw)] = \
                lambda xin=xin, w_hi=w_hi, w_lo=w_lo, kh=kh, kw=kw: ops.conv_gemm(xin, w_hi, w_lo, kh, kw, None, "relu")
        # the same layer at the three accumulator tile widths (tensor-pipe use per width)
        xin = torch.randn(1, 68, 120, 256, generator=g).to(dev)
        wt = (torch.randn(128, 5 * 256, generator=g) / (5 * 256) ** 0.5).to(dev)
        w_hi, w_lo, _ = ops.tc_weight([wt], lambda: (wt, None))
        for tn in (128, 64, 32):
            cases["conv_gemm_1x68x120_256->128_1x5_tile%d" % tn] = \
                lambda tn=tn: ops.conv_gemm(xin, w_hi, w_lo, 1, 5, None, "relu", tile_n=tn)
        # training side: weight gradient of the motion encoder's 3x3 256 -> 192 layer and its pre-passes
        xw = torch.randn(1, 68, 120, 256, generator=g).to(dev)
        gw = torch.randn(1, 68, 120, 192, generator=g).to(dev)
        gt = ops.transpose_split(gw)[:2]
        xt = ops.transpose_split(xw, n_shift=3)[:2]
        cases["transpose_split_x_3shifts_8160x256"] = lambda: ops.transpose_split(xw, n_shift=3)
        cases["conv_wgrad_1x68x120_256->192_3x3"] = lambda: ops.conv_wgrad(gt, xt, 1, 68, 120, 192, 256, 3, 3)
        cases["conv_wgrad_linear_2048x384->256"] = lambda: ops.conv_wgrad(
            ops.transpose_split(torch.randn(1, 1, 2048, 256, device=dev))[:2], ops.transpose_split(torch.randn(1, 1, 2048, 384, device=dev))[:2],
            1, 1, 2048, 256, 384, 1, 1)
        out = {}
        for name, fn in cases.items():
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                flush.add_(1.0)
                native.profile_begin()
                fn()
                prof = native.profile_end()
                ts.append({k: round(v["total_us"], 1) for k, v in prof.items()})
            out[name] = ts[-1] if ts else {}
            out[name + "/min"] = {k: min(t[k] for t in ts) for k in ts[0]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
