"""Per-shape CUDA-event timing of every C-ABI launch of one eager C2 forward (a diagnosis tool).

Every `native.call` of an eager forward is bracketed by events on its stream and keyed by
(entry point, shape arguments); prints per-key launches / avg us / GFLOP/s sorted by total
time and writes gpurun_out/conv_shapes.json.

    python scripts/trace_conv_shapes.py [--workload c2] [--passes 3]
"""
import argparse
import collections
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camliflow_b200 import native  # noqa: E402
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.engine import FlowEngine  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--passes", type=int, default=3)
    args = ap.parse_args()
    H, W, N, iters, B = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=iters)), seed=0)
    engine = FlowEngine(model, B, H, W, N, device=dev, use_graph=False)
    engine.load({k: v.pin_memory() for k, v in bench.synthetic_inputs(B, H, W, N, 0).items()})
    with torch.cuda.stream(engine.stream), torch.no_grad():
        engine._forward_static()
    torch.cuda.synchronize()

    recs = []
    real_call = native.call

    def traced(name, *a, algo_bytes=0, flops=0):
        ints = tuple(int(x.value) for x in a if isinstance(x, (ctypes.c_int, ctypes.c_int64)))
        st = torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        real_call(name, *a)
        e.record(st)
        recs.append((name, ints, s, e, algo_bytes, flops))

    native.call = traced
    with torch.cuda.stream(engine.stream), torch.no_grad():
        for _ in range(args.passes):
            engine._forward_static()
    torch.cuda.synchronize()
    native.call = real_call

    agg = collections.OrderedDict()
    for name, ints, s, e, by, fl in recs:
        if name == "camli_conv_gemm_fused":      # B,H,W,Cin,ldx,Cout,kh,kw,ldr,act,ldo,...
            key = "%s B%d %dx%d Cin%d->Cout%d %dx%d act%d" % (name, ints[0], ints[1], ints[2], ints[3], ints[5], ints[6], ints[7], ints[9])
        else:
            key = name + " " + ",".join(str(i) for i in ints[:8])
        r = agg.setdefault(key, [0, 0.0, by, fl])
        r[0] += 1
        r[1] += s.elapsed_time(e) * 1e3
    rows = []
    for key, (n, us, by, fl) in agg.items():
        rows.append({"key": key, "launches_per_pass": n / args.passes, "avg_us": us / n, "total_us_per_pass": us / args.passes,
                     "GFLOPs": fl / (us / n * 1e-6) / 1e9 if fl else 0.0, "GBps": by / (us / n * 1e-6) / 1e9 if by else 0.0,
                     "flops": fl, "bytes": by})
    rows.sort(key=lambda r: -r["total_us_per_pass"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "conv_shapes.json"), "w"), indent=1)
    tot = sum(r["total_us_per_pass"] for r in rows)
    print("total us per pass (sum over launches, two streams): %.0f" % tot)
    for r in rows[:70]:
        print("%7.1f us x %5.1f = %8.0f us  %8.0f GF/s %7.0f GB/s  %s" % (r["avg_us"], r["launches_per_pass"], r["total_us_per_pass"],
                                                                         r["GFLOPs"], r["GBps"], r["key"]))


if __name__ == "__main__":
    main()
