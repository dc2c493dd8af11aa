"""In-graph time per launch of camli_conv_gemm for a few representative layers (A/B of library variants: CAMLI_LIB_PATH)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200 import ops  # noqa: E402

SHAPES = [(1, 1, 2048, 128, 128, 1, 1, 0), (1, 1, 2048, 384, 128, 1, 1, 0), (1, 68, 120, 128, 128, 1, 1, 0),
          (1, 68, 120, 256, 192, 3, 3, 0), (1, 68, 120, 256, 256, 1, 5, 0), (1, 68, 120, 256, 128, 1, 5, 64),
          (4, 68, 120, 256, 192, 3, 3, 0), (4, 68, 120, 128, 128, 1, 1, 0), (2, 136, 240, 64, 64, 3, 3, 0)]
if os.environ.get("CG_TIME_SMALL"):       # small layers at the three tile widths (latency policy study)
    SHAPES = [(1, 1, 2048, 128, 128, 1, 1, t) for t in (128, 64, 32)] + [(1, 68, 120, 128, 128, 1, 1, t) for t in (128, 64, 32)] + \
             [(1, 1, 2048, 384, 256, 1, 1, t) for t in (128, 64, 32)] + [(1, 68, 120, 324, 256, 1, 1, t) for t in (128, 64)]


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    row = []
    for (B, H, W, Ci, Co, kh, kw, tile_n) in SHAPES:
        x = torch.randn(B, H, W, Ci, generator=g).to(dev)
        wt = (torch.randn(Co, kh * kw * Ci, generator=g) / (kh * kw * Ci) ** 0.5).to(dev)
        w_hi, w_lo, _ = ops.tc_weight([wt], lambda: (wt, None))
        out = torch.empty(B, H, W, Co, device=dev)
        fn = lambda: ops.conv_gemm(x, w_hi, w_lo, kh, kw, None, "relu", out=out, tile_n=tile_n)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n = 40
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(n):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            gr.replay()
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e3 / n)
        row.append(best)
    print(" ".join("%7.1f" % v for v in row))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--header":
        print(" ".join("%7s" % ("%dx%d.%d>%d" % (s[0], s[5] * s[6], s[3], s[4]))[:7] for s in SHAPES))
        print(" ".join("%7s" % ("r%d t%d" % (s[1] * s[2], s[7])) for s in SHAPES))
    main()
