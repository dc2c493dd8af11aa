"""Kernel timeline of one C2 forward (eager and CUDA-graph replay) through torch.profiler (CUPTI):
per-kernel totals, per-stream busy time and the wall span, written as JSON (+ the raw chrome
trace) under gpurun_out/.  A diagnosis tool: numbers taken under the profiler are never bench values.

    python scripts/trace_forward.py [--workload c2] [--out gpurun_out/trace]
"""
import argparse
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.engine import FlowEngine  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402


def summarise(trace_path):
    ev = json.load(open(trace_path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    if not ks:
        return {}
    t0 = min(e["ts"] for e in ks)
    t1 = max(e["ts"] + e["dur"] for e in ks)
    by_name = collections.defaultdict(lambda: [0, 0.0])
    by_stream = collections.defaultdict(float)
    for e in ks:
        by_name[e["name"][:90]][0] += 1
        by_name[e["name"][:90]][1] += e["dur"]
        by_stream[str(e.get("args", {}).get("stream"))] += e["dur"]
    # union of busy intervals over all streams (GPU idle = span - union)
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ks)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    top = sorted(by_name.items(), key=lambda kv: -kv[1][1])
    return {"span_us": t1 - t0, "busy_union_us": busy, "n_kernels": len(ks), "sum_us": sum(e["dur"] for e in ks),
            "by_stream_us": dict(by_stream), "kernels": [{"name": n, "n": c, "us": round(u, 1)} for n, (c, u) in top]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    H, W, N, iters, B = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    model = bench.build_model(args.workload)
    engine = FlowEngine(model, B, H, W, N, device=dev, use_graph=True)
    engine.load({k: v.pin_memory() for k, v in bench.synthetic_inputs(B, H, W, N, 0).items()})
    out = {}
    for mode in ("graph", "eager"):
        fn = engine.step if mode == "graph" else (lambda: engine._forward_static())
        with torch.cuda.stream(engine.stream):
            for _ in range(2):
                fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            with torch.cuda.stream(engine.stream):
                fn()
            torch.cuda.synchronize()
        path = "%s_%s.json" % (args.out, mode)
        prof.export_chrome_trace(path)
        out[mode] = summarise(path)
        if mode == "eager":
            os.remove(path)        # the eager trace carries every CPU op: large, and the graph one is what is judged
    json.dump(out, open(args.out + "_summary.json", "w"), indent=1)
    for mode, s in out.items():
        print(mode, {k: v for k, v in s.items() if k != "kernels"})
        for k in s.get("kernels", [])[:25]:
            print("   %6d %10.1f us  %s" % (k["n"], k["us"], k["name"]))


if __name__ == "__main__":
    main()
