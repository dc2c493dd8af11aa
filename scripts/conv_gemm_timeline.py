"""In-kernel timeline of camli_conv_gemm (CTA 0): SM-clock stamps of the producer / converter / MMA /
epilogue events, printed in cycles relative to kernel entry.  Also times back-to-back warm launches."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200 import native, ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    lib = native.lib()
    for (B, H, W, Ci, Co, kh, kw, tile_n) in [(1, 1, 2048, 384, 128, 1, 1, 0), (1, 1, 2048, 128, 128, 1, 1, 0),
                                              (1, 1, 2048, 384, 128, 1, 1, 128), (1, 68, 120, 256, 192, 3, 3, 0),
                                              (1, 68, 120, 256, 128, 1, 5, 0), (1, 68, 120, 256, 128, 1, 5, 64),
                                              (1, 68, 120, 256, 126, 3, 3, 64), (1, 68, 120, 128, 128, 1, 1, 0),
                                              (1, 1, 2048, 4, 32, 1, 1, 0), (4, 68, 120, 256, 192, 3, 3, 0),
                                              (4, 68, 120, 128, 128, 1, 1, 0), (2, 136, 240, 64, 64, 3, 3, 0)]:
        x = torch.randn(B, H, W, Ci, generator=g).to(dev)
        wt = (torch.randn(Co, kh * kw * Ci, generator=g) / (kh * kw * Ci) ** 0.5).to(dev)
        w_hi, w_lo, _ = ops.tc_weight([wt], lambda: (wt, None))
        out = torch.empty(B, H, W, Co, device=dev)
        fn = lambda: ops.conv_gemm(x, w_hi, w_lo, kh, kw, None, "relu", out=out, tile_n=tile_n)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # accuracy against an fp64 convolution of the same operands
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.view(Co, kh, kw, Ci).permute(0, 3, 1, 2).double(),
                                         padding=(kh // 2, kw // 2)).relu().permute(0, 2, 3, 1)
        err = float((out.double() - ref).abs().max() / ref.pow(2).mean().sqrt())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        per = s.elapsed_time(e) * 1e3 / n
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(n):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        s.record()
        gr.replay()
        e.record()
        torch.cuda.synchronize()
        per_g = s.elapsed_time(e) * 1e3 / n
        buf = torch.zeros(128, dtype=torch.int64, device=dev)
        lib.camli_conv_gemm_set_timeline(ctypes.c_void_p(buf.data_ptr()))
        fn()
        torch.cuda.synchronize()
        lib.camli_conv_gemm_set_timeline(ctypes.c_void_p(0))
        t = buf.cpu().tolist()
        t0 = t[0]
        rel = lambda i: (t[i] - t0) if t[i] else None  # noqa: E731
        print("== %s tile_n=%d: warm back-to-back %.1f us/launch eager, %.1f us/launch in a graph, max err / rms %.2e" %
              ((B, H, W, Ci, Co, kh, kw), tile_n, per, per_g, err))
        print("   setup done %s | cfull committed %s | cfull seen %s | tile stored %s | exit %s  (cycles)" %
              (rel(1), rel(8), rel(9), rel(10), rel(11)))
        print("   epilogue: corr loaded %s, bias/residual done %s | producer: armed %s, first TMA issued %s" %
              (rel(12), rel(13), rel(14), rel(15)))
        print("   TMA issued   ", [rel(16 + i) for i in range(12)])
        print("   data landed  ", [rel(32 + i) for i in range(12)])
        print("   split done   ", [rel(64 + i) for i in range(12)])
        print("   MMAs issued  ", [rel(48 + i) for i in range(12)])
        print("   chunk seen   ", [rel(80 + i) for i in range(8)])


if __name__ == "__main__":
    main()
