"""Turns the text outputs of scripts/gpu_profile.sh (gpurun_out/) into the tracked evidence under profiles/:
  r<round>_bench.json            the bench line (as printed)
  r<round>_launches.md/.json     ncu launch list of the same bench command: per-kernel counts, time, share of the step
  r<round>_kernels.md/.json      ncu --set full of the hot kernels: duration, DRAM bytes, achieved bandwidth / pipe use
    python scripts/summarise_profiles.py [--round 1]
"""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void |at::native::", "", name)
    return re.sub(r"\(.*", "", name)[:90]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r[-1].replace(",", ""))
        us = v / 1000.0 if r[-2] in ("ns", "nsecond") else (v * 1000.0 if r[-2] in ("ms", "msecond") else v)
        agg[short(r[4])][0] += 1
        agg[short(r[4])][1] += us
    total = sum(v[1] for v in agg.values())
    library = ("at::", "cutlass", "cudnn", "sm80_", "sm90_", "sm100_", "convolve_", "cublas", "implicit_convolve", "nhwc", "vectorized_",
               "elementwise_", "CatArray", "max_pool", "void cudnn", "void cutlass")
    table = [{"kernel": k, "launches": c, "total_us": round(u, 1), "share": round(u / total, 4),
              "hand_written": not k.startswith(library)}
             for k, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump({"command": "CAMLI_PROFILER_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum "
                          "--clock-control none python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline",
               "note": "cold-cache, serialised per-launch times: compare SHARES, not absolutes", "n_launches": len(rows),
               "total_us": round(total, 1), "kernels": table}, open(os.path.join(PROF, tag + "_launches.json"), "w"), indent=1)
    with open(os.path.join(PROF, tag + "_launches.md"), "w") as f:
        f.write("# ncu launch list of one C2 step (%d launches, %.1f ms serialised, cold cache)\n\n" % (len(rows), total / 1e3))
        f.write("| kernel | launches | total µs | share | ours |\n|---|---:|---:|---:|---|\n")
        for t in table[:45]:
            f.write("| `%s` | %d | %.1f | %.1f%% | %s |\n" % (t["kernel"], t["launches"], t["total_us"], 100 * t["share"],
                                                            "yes" if t["hand_written"] else ""))
        hw = sum(t["share"] for t in table if t["hand_written"])
        f.write("\nHand-written kernels: %.1f%% of the serialised step time.\n" % (100 * hw))


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def kernels(tag, peaks):
    path = os.path.join(OUT, "prof_kernels_raw.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {m: hdr.index(m) for m in METRICS if m in hdr}
    name_i = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        rec = {"kernel": short(r[name_i])}
        for m, i in col.items():
            try:
                rec[m] = float(r[i].replace(",", ""))
            except ValueError:
                rec[m] = r[i]
            rec.setdefault("_units", {})[m] = units[i]
        u = rec["_units"]
        t_us = rec["gpu__time_duration.sum"] * {"us": 1, "usecond": 1, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(
            u["gpu__time_duration.sum"], 1)
        scale = lambda m: {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u.get(m, "byte"), 1)   # noqa: E731
        traffic = rec.get("dram__bytes_read.sum", 0) * scale("dram__bytes_read.sum") + \
            rec.get("dram__bytes_write.sum", 0) * scale("dram__bytes_write.sum")
        rec.update({"duration_us": round(t_us, 2), "dram_traffic_MB": round(traffic / 1e6, 2),
                    "dram_GBps": round(traffic / (t_us * 1e-6) / 1e9, 1),
                    "dram_frac_of_measured_peak": round(traffic / (t_us * 1e-6) / 1e9 / peaks["hbm_gbs"], 3)})
        del rec["_units"]
        out.append(rec)
    json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:<hot kernels> python "
                          "scripts/profile_kernels.py --iters 1", "hbm_peak_GBps_measured": peaks["hbm_gbs"], "kernels": out},
              open(os.path.join(PROF, tag + "_kernels.json"), "w"), indent=1)
    with open(os.path.join(PROF, tag + "_kernels.md"), "w") as f:
        f.write("# ncu --set full, hot kernels at C2 sizes (cold cache, one launch each)\n\n")
        f.write("| kernel | grid x block | regs | µs | DRAM MB (r+w) | DRAM GB/s | of measured %.0f GB/s | warps active %% | issue active %% | tensor pipe %% |\n"
                % peaks["hbm_gbs"])
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in out:
            f.write("| `%s` | %d x %d | %d | %.1f | %.1f | %.0f | %.2f | %.0f | %.0f | %.1f |\n" % (
                r["kernel"][:60], r.get("launch__grid_size", 0), r.get("launch__block_size", 0),
                r.get("launch__registers_per_thread", 0), r["duration_us"], r["dram_traffic_MB"], r["dram_GBps"],
                r["dram_frac_of_measured_peak"], r.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0),
                r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
                r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) or 0))


ENTRY_POINTS = {          # ncu kernel name prefix -> C-ABI entry point (the key bench.py's roofline block looks up)
    "corr2d_lookup_kernel": "camli_corr2d_lookup",
    "dw_gather_max": "camli_pointconv_dw_gather_max",
    "conv_gemm_tf32x3_kernel": "camli_conv_gemm_strided",
    "conv_wgrad_tf32x3_kernel": "camli_conv_wgrad",
    "allpairs_tf32x3_kernel": "camli_allpairs_correlation",
    "fps_cluster_async_kernel": "camli_furthest_point_sampling",
    "fps_pruned_kernel": "camli_furthest_point_sampling",
    "convex_upsample_kernel": "camli_convex_upsample",
    "corr3d_lookup_kernel": "camli_corr3d_lookup",
}


def ncu_summary(tag, workload="c2"):
    """profiles/<tag>_ncu_summary.json: per entry point, the ncu figures bench.py may quote beside its live timings
    (largest-traffic launch of each kernel family in <tag>_kernels.json)."""
    import subprocess
    path = os.path.join(PROF, tag + "_kernels.json")
    if not os.path.exists(path):
        return
    doc = json.load(open(path))
    best = {}
    for r in doc["kernels"]:
        for prefix, entry in ENTRY_POINTS.items():
            if r["kernel"].startswith(prefix):
                cur = best.get(entry)
                if cur is None or r["dram_traffic_MB"] > cur["dram_bytes"] / 1e6:
                    best[entry] = {"kernel": r["kernel"], "duration_us": r["duration_us"], "dram_bytes": r["dram_traffic_MB"] * 1e6,
                                   "dram_GBps": r["dram_GBps"], "dram_frac_of_measured_peak": r["dram_frac_of_measured_peak"],
                                   "tensor_pipe_pct": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                   "grid": r.get("launch__grid_size")}
    try:
        sha = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, text=True).strip()
    except Exception:
        sha = None
    json.dump({"workload": workload, "commit": sha, "source": tag + "_kernels.json (ncu --set full, scripts/profile_kernels.py, C2 sizes, "
               "cold cache)", "kernels": best}, open(os.path.join(PROF, tag + "_ncu_summary.json"), "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", type=int, default=1)
    ap.add_argument("--suffix", default="")
    args = ap.parse_args()
    tag = "r%d%s" % (args.round, args.suffix)
    os.makedirs(PROF, exist_ok=True)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    bench = os.path.join(OUT, "bench.json")
    if os.path.exists(bench) and os.path.getsize(bench):
        line = json.loads(open(bench).read().strip().splitlines()[-1])
        json.dump(line, open(os.path.join(PROF, tag + "_bench.json"), "w"), indent=1)
    launches(tag)
    kernels(tag, peaks)
    ncu_summary(tag)
    for extra in ("clocks.csv",):
        src = os.path.join(OUT, extra)
        if os.path.exists(src):
            lines = open(src).read().strip().splitlines()
            open(os.path.join(PROF, tag + "_clocks_summary.txt"), "w").write(
                "%d samples during the bench run; header + first/last rows:\n%s\n%s\n%s\n" % (len(lines) - 1, lines[0], lines[1], lines[-1]))
    print("wrote", sorted(f for f in os.listdir(PROF) if f.startswith(tag)))


if __name__ == "__main__":
    main()
