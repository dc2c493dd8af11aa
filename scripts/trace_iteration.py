"""Timeline of ONE refinement iteration of the captured C2 forward, from the chrome trace scripts/trace_forward.py
writes (gpurun_out/trace_graph.json): every kernel between the 6th and the 7th corr2d_lookup launch with its start
offset, duration and stream, plus the pre-loop milestones.  A diagnosis tool (numbers under CUPTI are not bench values).

    python scripts/trace_iteration.py [gpurun_out/trace_graph.json]
"""
import json
import sys


def short(n):
    return n.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")[:64]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_graph.json"
    ev = json.load(open(path))["traceEvents"]
    ks = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e), key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in ks) - t0
    look = [e for e in ks if "corr2d_lookup_kernel" in e["name"]]
    print("span %.1f us, %d kernels, %d lookups" % (end, len(ks), len(look)))
    for key in ("fps_", "stem_conv", "allpairs", "corr2d_pool"):
        for e in ks:
            if key in e["name"]:
                print("  %-14s start %8.1f end %8.1f (stream %s)" % (key, e["ts"] - t0, e["ts"] + e["dur"] - t0, e["args"]["stream"]))
    if len(look) >= 8:
        print("  loop: first lookup %.1f, iteration period %.1f us, after the last lookup %.1f us" %
              (look[0]["ts"] - t0, (look[-1]["ts"] - look[0]["ts"]) / (len(look) - 1), end - (look[-1]["ts"] - t0)))
        a, b = look[5]["ts"], look[6]["ts"]
        by_stream = {}
        for e in ks:
            if a <= e["ts"] < b:
                by_stream[e["args"]["stream"]] = by_stream.get(e["args"]["stream"], 0.0) + e["dur"]
                print("  %7.1f +%6.1f s%-4s %s" % (e["ts"] - a, e["dur"], e["args"]["stream"], short(e["name"])))
        print("  busy per stream in this iteration:", {k: round(v, 1) for k, v in by_stream.items()}, "period %.1f" % (b - a))


if __name__ == "__main__":
    main()
