#!/usr/bin/env python
"""Benchmark of the CamLiRAFT hot path (BASELINE.json metric: frame-pairs/s, 960x540 RGB +
8192 points, 12 GRU iterations, batch 1 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N>1); the batch of frame pairs is sharded over the ranks
with no data-path collective ("scaling": "weak").  Prints ONE JSON line on rank 0.

A "step" streams `--pairs-per-step` (default 8) independent frame pairs through the engine(s); every forward is a
batch-1 CUDA graph (BASELINE config 2), so a default run times >= 160 pairs (>= 1.5 s) instead of a fraction of a second.

* value        : pairs/s with the inputs resident in HBM (device-timed, CUDA events, max over ranks), throughput mode:
                 3 batch-1 graphs in flight over shared weights (EnginePool)
* latency      : one graph at a time: ms per frame pair (L2 flushed before every pair) -- the batch-1 latency
* e2e          : throughput end to end from HOST tensors (pinned H2D + forward + D2H for every pair) through the public
                 API EnginePool.pipelined; the single-engine pipelined and the blocking call-per-pair figures beside it
* roofline     : the hand-written kernel with the largest share of the step, timed per launch with CUDA events on its
                 stream (live); `traffic` from the committed ncu capture under profiles/ when it matches the workload
* training     : BASELINE config 5 in the same run (a few steps): fwd + losses + bwd + DDP gradient all-reduce over
                 NCCL + AdamW, so the one collective of this project is on the driver's clock at every N
* cpu_baseline : the reference algorithm's CPU path (oracle port, fallback index semantics)
                 timed on this host's cores on a bounded sample (rank 0, N=1 only)
* --impl reference : only the CPU path (the reference cannot travel to the GPU box; its
                 restatement oracle/camliraft_oracle.py, pinned against it, is what runs)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (H, W, N points, GRU iterations, pairs per forward on one GPU)
    "c2": (540, 960, 8192, 12, 1),
    "small": (160, 224, 8192, 3, 1),
    # BASELINE config 3 (`--workload c3`): CamLiPWC (5-level PWC cost volume), batch 4
    "c3": (540, 960, 8192, 0, 4),
    # BASELINE config 4 (`--workload c4`): 32 iterations, 32 pairs over 8 GPUs = 4 pairs per GPU
    "c4": (540, 960, 8192, 32, 4),
    # BASELINE config 5 (secondary; `--workload c5`): training step, n_iters_train = 10, 2 pairs per GPU
    "c5": (540, 960, 8192, 10, 2),
    "c5small": (160, 224, 8192, 3, 1),
}
TRAIN_WORKLOADS = ("c5", "c5small")
METRIC = "CamLiRAFT frame-pairs/sec 960x540+8192pts"


def workload_name(workload):
    """The `config.workload` string -- identical in both arms (ours / --impl reference)."""
    H, W, N, iters, B = WORKLOADS[workload]
    if workload == "c3":
        return "c3: CamLiPWC fusion %dx%d RGB + %d pts, batch %d per forward" % (W, H, N, B)
    if workload in TRAIN_WORKLOADS:
        return "%s: CamLiRAFT training step (fwd + sequence losses + bwd + AdamW) %dx%d RGB + %d pts, %d iters, batch %d per GPU" \
            % (workload, W, H, N, iters, B)
    return "%s: CamLiRAFT fusion %dx%d RGB + %d pts, %d iters, batch %d per forward" % (workload, W, H, N, iters, B)


def build_model(workload):
    from camliflow_b200.init import seed_module_
    iters = WORKLOADS[workload][3]
    if workload == "c3":
        from camliflow_b200.camlipwc import CamLiPWC
        from camliflow_b200.config import camlipwc_config
        return seed_module_(CamLiPWC(camlipwc_config()), seed=0)
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    return seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=iters, n_iters_train=iters)), seed=0)


def synthetic_inputs(B, H, W, N, seed):
    """SURVEY 8(d) generator (identical to oracle.camliraft_oracle.synthetic_inputs; tests check)."""
    g = torch.Generator().manual_seed(seed)
    f, cx, cy = 1050.0, (W - 1) / 2.0, (H - 1) / 2.0
    images = torch.randint(0, 256, (B, 6, H, W), generator=g).float()
    u = torch.rand((B, N), generator=g) * (W - 1)
    v = torch.rand((B, N), generator=g) * (H - 1)
    z = torch.rand((B, N), generator=g) * 30.0 + 5.0
    pc1 = torch.stack([(u - cx) * z / f, (v - cy) * z / f, z], 1)
    pc2 = pc1 + torch.randn(pc1.shape, generator=g) * 0.05
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    pc2 = torch.gather(pc2, 2, perm[:, None, :].expand(B, 3, N))
    return {"images": images, "pcs": torch.cat([pc1, pc2], 1), "intrinsics": torch.tensor([[f, cx, cy]]).repeat(B, 1)}


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU path
def run_cpu_reference(workload, steps, warmup, seed=0):
    """Times the reference algorithm's CPU path: oracle port, torch fallbacks for FPS / k-NN (what
    models/csrc/wrapper.py does for CPU tensors), all host threads, eval mode.  Like our arm it materialises the
    final prediction only (the reference also up-samples the 11 intermediate ones, which nothing reads at
    inference: leaving that out favours the baseline, so the ratio is like for like)."""
    from oracle import camliraft_oracle as co
    H, W, N, iters, B = WORKLOADS[workload]
    if workload == "c3":
        return None            # no CPU restatement of CamLiPWC is kept (the C2 metric is the headline)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = co.make_params(co.param_spec("camliraft"), seed=0)
    inp = co.synthetic_inputs(B, H, W, N, seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        co.camliraft_forward(P, inp["images"], inp["pcs"], inp["intrinsics"], n_iters=iters, index_impl="fallback",
                             all_iters=False)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": B / sec, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d full forward(s) of %d pair(s), %dx%d + %d pts, %d iters, final prediction only, after %d warm-up"
                      % (steps, B, W, H, N, iters, warmup), "sec_per_step": sec}


# ---------------------------------------------------------------------------------- GPU path
def l2_flush(buf):
    buf.add_(1.0)


def shard_seed(rank, base_seed=0):
    """Every rank owns its own frame pairs (weak scaling over independent pairs): the generator seed
    of rank r's shard."""
    return base_seed + rank


def max_over_ranks(ms, world, device=None):
    """Device-timed milliseconds, reduced with MAX over the ranks (identity for one rank)."""
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pairs_per_step(B, world, forwards=1):
    return B * world * forwards


TENSOR_KERNELS = ("camli_conv_gemm_strided", "camli_conv_gemm_fused", "camli_conv_gemm", "camli_allpairs_correlation")
NAMED_HBM_KERNELS = (("camli_corr2d_lookup", "corr_lookup"), ("camli_pointconv_dw_gather_max", "knn_gather"))


def committed_ncu(workload):
    """Per-kernel ncu figures (duration, DRAM bytes) from the newest committed profiles/r*_kernels.json whose
    `workload` field matches, or {} -- never constants in code."""
    import glob
    best = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json"))):
        try:
            doc = json.load(open(path))
        except (OSError, ValueError):
            continue
        if doc.get("workload") == workload:
            best = dict(doc.get("kernels", {}), source=os.path.relpath(path, ROOT), commit=doc.get("commit"))
    return best


def hold_gpu(device, ms):
    """Keep the current stream busy for ~`ms` so that the launches enqueued next queue up behind it: an event pair around a
    launch then brackets the kernel alone, not the host's launch latency of an idle GPU (5-10 us per eager launch)."""
    khz = torch.cuda.get_device_properties(device).clock_rate or 1900000
    torch.cuda._sleep(int(ms * khz))


def same_size_copy_us(n_bytes, device, flush):
    """Yardstick for the small HBM-bound kernels: a plain device-to-device copy moving the same number of bytes (half read,
    half written), timed like them -- queued behind a busy GPU, one CUDA-event pair per launch, L2 flushed before each, minimum
    of 10.  At a few tens of MB the DRAM ramp dominates: the copy itself stays far below the large-copy peak of
    MEASURED_PEAKS.json."""
    n = max(1, int(n_bytes) // 8)
    src, dst = torch.empty(n, dtype=torch.float32, device=device), torch.empty(n, dtype=torch.float32, device=device)
    pairs = []
    hold_gpu(device, 2.0)
    for i in range(13):
        flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dst.copy_(src)
        e.record()
        pairs.append((s, e))
    torch.cuda.synchronize(device)
    return min(s.elapsed_time(e) * 1e3 for s, e in pairs[3:])


def gemm_yardstick(shape, device):
    """Yardstick for the tensor-core kernel: cuBLAS GEMMs of the largest layer's (M, N, K) on the same box, timed back to back
    with CUDA events -- single-pass TF32 (the tensor-core rate the library reaches at this SIZE, at 10-bit mantissa accuracy)
    and fp32 SGEMM (the library kernel of the same accuracy class as the 3xTF32 kernel)."""
    M, N, K = shape
    a = torch.randn(M, K, device=device)
    b = torch.randn(K, N, device=device)
    out, old = {}, torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("cublas_tf32_1pass", True), ("cublas_fp32", False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(3):
                torch.matmul(a, b)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                torch.matmul(a, b)
            e.record()
            torch.cuda.synchronize(device)
            us = s.elapsed_time(e) * 100.0
            out[name] = {"us": us, "TFLOPs": 2.0 * M * N * K / (us * 1e-6) / 1e12}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def roofline_block(prof, workload, copy_yardstick=None, gemm_yardstick_fn=None):
    """Roofline record of the hand-written kernel with the largest share of the profiled (eager) step, from live
    CUDA-event timings.  Tensor-core kernels (3xTF32 implicit GEMM): achieved = ISSUED tf32 flops (3 products per
    fp32 product) per launch / duration against the tf32 dense peak (half the measured bf16 peak: same tensor
    pipe, K = 8 instead of 16 per instruction); the fp32-equivalent rate is beside it.  HBM-bound kernels:
    algorithmic bytes per launch / duration against the measured copy bandwidth."""
    if not prof:
        return None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, bf16, src = 6650.0, 1590.0, "fallback"
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        hbm, bf16, src = float(peaks["hbm_gbs"]), float(peaks.get("bf16_tflops", 1590.0)), "measured"
    ncu = committed_ncu(workload)
    per = {n: {"avg_us": r["avg_us"], "launches": r["launches"], "total_us": r["total_us"], "by_size": r["by_size"][:6],
               "GBps": r["bytes"] / (r["avg_us"] * 1e-6) / 1e9, "GFLOPs": r["flops"] / (r["avg_us"] * 1e-6) / 1e9}
           for n, r in prof.items()}
    total = sum(r["total_us"] for r in prof.values())
    top = max(prof, key=lambda n: prof[n]["total_us"])
    rec = prof[top]
    if top in TENSOR_KERNELS:
        fp32_eq = rec["flops"] / (rec["avg_us"] * 1e-6) / 1e12
        out = {"kernel": top, "bound": "tensor", "achieved": 3.0 * fp32_eq, "peak": bf16 / 2.0,
               "peak_source": src + " dense bf16 / 2 (tf32 runs at half the bf16 rate)", "unit": "TFLOP/s",
               "frac": 3.0 * fp32_eq / (bf16 / 2.0), "fp32_equivalent_TFLOPs": fp32_eq,
               "note": "achieved = issued tf32 flops: 3 tf32 products per fp32 product (3xTF32); averaged over every launch "
                       "of the step, from 16-CTA linear layers to 3x3 convolutions",
               "algorithmic_flops_per_launch": rec["flops"]}
        big = max(rec["by_size"], key=lambda g: g["flops"])          # the largest layer of the step
        out["largest"] = {"algorithmic_flops": big["flops"], "launches": big["launches"], "avg_us": big["avg_us"],
                          "fp32_equivalent_TFLOPs": big["flops"] / (big["avg_us"] * 1e-6) / 1e12,
                          "frac": 3.0 * big["flops"] / (big["avg_us"] * 1e-6) / 1e12 / (bf16 / 2.0)}
        if gemm_yardstick_fn is not None and big.get("shape"):
            out["largest"]["gemm_MNK"] = list(big["shape"])
            out["largest"]["same_shape_library_gemm"] = dict(
                gemm_yardstick_fn(big["shape"]),
                note="library GEMMs of the same M, N, K on this box: the hand-written kernel computes an fp32-accurate result "
                     "(error ~2e-6, 3 tf32 products per fp32 product); cublas_fp32 is the library kernel of that accuracy class, "
                     "cublas_tf32_1pass the single-product tensor-core rate the library reaches at this size")
    else:
        achieved = rec["bytes"] / (rec["avg_us"] * 1e-6) / 1e9
        out = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm, "peak_source": src, "unit": "GB/s",
               "frac": achieved / hbm, "algorithmic_bytes_per_launch": rec["bytes"]}
    k = ncu.get(top)
    out.update({"traffic": k.get("dram_bytes") if k else None, "traffic_source": ncu.get("source") if k else None,
                "avg_us": rec["avg_us"], "launches": rec["launches"], "share_of_step": rec["total_us"] / total, "all": per})
    # the two HBM-bound kernels the north star names: live CUDA-event figures of this run (cold launches of an eager
    # pass, launch overhead included) and, when a committed ncu capture of this workload exists, its figures
    for name, key in NAMED_HBM_KERNELS:
        if name in per:
            out[key] = {"achieved_GBps": per[name]["GBps"], "frac_of_hbm_peak": per[name]["GBps"] / hbm,
                        "avg_us": per[name]["avg_us"], "launches": per[name]["launches"],
                        "algorithmic_bytes_per_launch": prof[name]["bytes"]}
            big = prof[name]["by_size"][0]               # the largest problem size of the step (the SURVEY 8(d) example shapes)
            out[key]["largest"] = {"algorithmic_bytes": big["bytes"], "launches": big["launches"], "avg_us": big["avg_us"],
                                   "min_us": big["min_us"], "achieved_GBps": big["bytes"] / (big["avg_us"] * 1e-6) / 1e9,
                                   "frac_of_hbm_peak": big["bytes"] / (big["avg_us"] * 1e-6) / 1e9 / hbm}
            if copy_yardstick is not None:
                cu = copy_yardstick(big["bytes"])
                out[key]["largest"].update({"same_size_copy_us": cu, "frac_of_same_size_copy": cu / big["avg_us"]})
                cu = copy_yardstick(prof[name]["bytes"])
                out[key].update({"same_size_copy_us": cu, "same_size_copy_frac_of_hbm_peak": prof[name]["bytes"] / (cu * 1e-6) / 1e9 / hbm,
                                 "frac_of_same_size_copy": cu / per[name]["avg_us"],
                                 "note": "same_size_copy = torch device copy moving the same bytes, timed the same way: what a "
                                         "perfect streaming kernel reaches at this size"})
            k = ncu.get(name)
            if k:
                out[key]["ncu"] = dict(k, source=ncu.get("source"), commit=ncu.get("commit"))
    return out


def pin_rank_to_cores(local_rank, world):
    """One process per GPU: give each rank its own slice of the host cores this job may use, so that the 8 feeding threads of
    an 8-GPU run (pinned H2D staging, graph launches) do not migrate over -- and evict each other from -- the same cores.
    Returns the cores taken (None: single process or no affinity API)."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // world
        if per < 1:
            return None
        mine = cpus[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(per, 8)))
        return mine
    except OSError:
        return None


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from camliflow_b200 import ops
    from camliflow_b200.engine import EnginePool, FlowEngine

    cores = pin_rank_to_cores(local_rank, world)
    H, W, N, iters, B = WORKLOADS[args.workload]
    R = args.pairs_per_step                        # forwards per step
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    strict = args.conv_precision == "fp32"
    torch.backends.cudnn.allow_tf32 = not strict
    ops.SINGLE_PASS_INFERENCE = not strict               # tf32: the dense kernel issues one tf32 product instead of three
    torch.backends.cuda.matmul.allow_tf32 = False      # GEMMs (1x1 layers, Linear) are always fp32
    model = build_model(args.workload)
    # the single engine runs one graph at a time (latency, blocking calls): 64-wide tiles for the half-empty layers; the pool's
    # engines share the GPU and keep the throughput tiling
    engine = FlowEngine(model, B, H, W, N, device=dev, use_graph=not args.no_graph, tile_policy="latency")
    inputs = synthetic_inputs(B, H, W, N, seed=shard_seed(rank))   # per-rank shard of the batch of pairs
    pinned = {k: v.pin_memory() for k, v in inputs.items()}
    flush = torch.zeros(192 * 1024 * 1024 // 4, device=dev)   # 192 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, warm):
        """n forwards, one CUDA-event pair around each (L2 flushed before, outside the pair); returns total ms."""
        for _ in range(warm):
            fn()
        barrier()
        evs = []
        with torch.cuda.stream(engine.stream):
            for _ in range(n):
                l2_flush(flush)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(engine.stream)
                fn()
                e.record(engine.stream)
                evs.append((s, e))
        barrier()
        return max_over_ranks(sum(s.elapsed_time(e) for s, e in evs), world, dev)

    engine.load(pinned)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if os.environ.get("CAMLI_PROFILER_RANGE"):       # ncu --profile-from-start off: capture the timed steps only
        torch.cuda.profiler.start()
    n_lat = max(args.steps, 10)
    ms_lat = timed(engine.step, n_lat, args.warmup)                      # one graph at a time: latency of a pair
    if os.environ.get("CAMLI_PROFILER_RANGE"):
        torch.cuda.profiler.stop()
    ms_sync = timed(lambda: engine(pinned), n_lat, args.warmup)          # blocking public call per pair

    n_fwd, n_warm = args.steps * R, args.warmup * R

    # single-engine throughput mode: H2D of pair i+1 and D2H of pair i-1 overlap the compute of pair i
    def piped(n):
        for out in engine.pipelined(pinned for _ in range(n)):
            pass
    piped(n_warm)
    barrier()
    t0 = time.perf_counter()
    piped(n_fwd)
    barrier()
    ms_pipe = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)

    # serving-style concurrency: several engines (CUDA graphs over the same weights) fed round-robin, so the
    # latency-bound phases of one pair overlap another pair's work; every forward is still batch B
    n_eng = max(1, args.concurrent)
    pool = EnginePool(model, n_eng, B, H, W, N, device=dev, use_graph=not args.no_graph) if n_eng > 1 else None
    if pool is not None:
        def pooled(n):
            for out in pool.pipelined(pinned for _ in range(n)):
                pass
        pooled(n_warm + n_eng)
        barrier()
        t0 = time.perf_counter()
        pooled(n_fwd)
        barrier()
        ms_pool = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
        # the same pool with the inputs resident in HBM (no host traffic), device-timed: one start event every engine
        # stream waits on, one end event after all of them joined a control stream
        for eng in pool.engines:
            eng.load(pinned)
        ctl = torch.cuda.Stream(dev)

        def resident(n, timed_pair=None):
            if timed_pair is not None:
                timed_pair[0].record(ctl)
                for eng in pool.engines:
                    eng.stream.wait_event(timed_pair[0])
            for i in range(n):
                pool.engines[i % n_eng].step()
            for eng in pool.engines:
                ctl.wait_stream(eng.stream)
            if timed_pair is not None:
                timed_pair[1].record(ctl)
        resident(n_warm)
        barrier()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        resident(n_fwd, ev)
        barrier()
        ms_dev = max_over_ranks(ev[0].elapsed_time(ev[1]), world, dev)
        ms_e2e, e2e_mode = ms_pool, ("EnginePool.pipelined: %d CUDA graphs of batch %d in flight, round-robin; each pair pays its "
                                     "own pinned H2D + D2H (host wall clock)" % (n_eng, B))
    else:
        ms_dev = ms_lat / n_lat * n_fwd
        ms_e2e, e2e_mode = ms_pipe, "FlowEngine.pipelined: pinned H2D of pair i+1 and D2H of pair i-1 overlap the forward of pair i"

    clocks = sampler.stop() if rank == 0 else None      # sampled over every timed region above

    # per-launch timing of every hand-written kernel (eager pass, events on the launch stream)
    roofline = None
    if rank == 0:
        # single stream for this pass: with the image and the point branch overlapped, an event pair around a launch
        # would also time the other branch's kernels that hold the SMs
        core = getattr(model, "core", None)
        overlapped = getattr(core, "two_streams", None)
        if overlapped is not None:
            core.two_streams = False
        timing = "cuda-graph event nodes"
        try:
            # the single-stream forward captured in a CUDA graph of its own, a timing-event record node on either side of
            # every launch: the durations of the replay are those of the kernels as they run in production (back to back,
            # inputs in L2 or not exactly as there), free of the 5-10 us an eager launch waits on an idle GPU
            with torch.cuda.stream(engine.stream), torch.no_grad():
                engine._forward_static()
                engine.stream.synchronize()
                ops.profile_begin()
                pg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(pg, stream=engine.stream):
                    engine._forward_static()
                for _ in range(3):
                    l2_flush(flush)
                    pg.replay()
            engine.stream.synchronize()
            prof = ops.profile_end()
            del pg
        except Exception as exc:       # (older drivers: no timing through event nodes) -> eager pass behind a held stream
            print("bench: graph-event profile pass failed (%s); eager pass" % exc, file=sys.stderr)
            timing = "eager launches queued behind a held stream"
            try:
                ops.profile_end()
            except Exception:
                pass
            torch.cuda.synchronize(dev)
            ops.profile_begin()
            with torch.cuda.stream(engine.stream), torch.no_grad():
                for _ in range(2):
                    l2_flush(flush)
                    hold_gpu(dev, 60.0)
                    engine._forward_static()
            engine.stream.synchronize()
            prof = ops.profile_end()
        roofline = roofline_block(prof, args.workload, lambda nb: same_size_copy_us(nb, dev, flush),
                                  lambda shp: gemm_yardstick(shp, dev))
        roofline["timing"] = timing
        if overlapped is not None:
            core.two_streams = overlapped

    h2d, d2h = engine.io_bytes()
    pairs = pairs_per_step(B, world, R)              # frame pairs per step over all ranks
    all_pairs = pairs * args.steps
    line = {
        "metric": METRIC if args.workload != "c3" else "CamLiPWC frame-pairs/sec 960x540+8192pts",
        "value": all_pairs / (ms_dev / 1e3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if strict else "tf32 (reduced precision: not the headline mode)",
        "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "pairs_per_step": pairs, "forwards_per_step_per_gpu": R,
                   "engines_in_flight": n_eng, "cuda_graph": engine.graph is not None,
                   "host_cores_per_rank": None if cores is None else len(cores),
                   "l2": "value / e2e: %d CUDA graphs in flight, per-engine working set (355 MB volume pyramid + activations) "
                         "larger than L2, e2e inputs re-copied for every pair; latency: 192 MiB flush write before every "
                         "timed pair (outside the event pair)" % n_eng,
                   "conv_precision": ("fp32 (3xTF32 tensor-core kernels, fp32-accurate: the mode the parity tests run in)" if strict else
                                      "tf32 -- NOT the parity mode: one tf32 product per element in the dense kernels, as torch's "
                                      "default allow_tf32 gives the reference's cuDNN convolutions"),
                   "intermediate_predictions": False},
        # batch-1 latency (BASELINE config 2 is batch 1): one CUDA graph at a time
        "latency": {"ms_per_pair": ms_lat / n_lat / B, "pairs_per_s": B * world * n_lat / (ms_lat / 1e3), "pairs_timed": n_lat * B,
                    "note": "one CUDA graph at a time (tile policy 'latency': the narrowest tile that keeps a layer within one wave "
                            "of CTAs), device-timed per pair, L2 flushed before each"},
        "e2e": {"value": all_pairs / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d * R, "d2h_bytes_per_step": d2h * R, "mode": e2e_mode, "engines_in_flight": n_eng,
                "single_engine_pipelined": {"value": all_pairs / (ms_pipe / 1e3), "unit": "pairs/s",
                                            "mode": "FlowEngine.pipelined: pinned H2D of pair i+1 and D2H of pair i-1 overlap "
                                                    "the forward of pair i"},
                "synchronous": {"value": B * world * n_lat / (ms_sync / 1e3), "unit": "pairs/s", "ms_per_pair": ms_sync / n_lat / B,
                                "mode": "FlowEngine.__call__ per pair: pinned H2D -> forward -> D2H -> host sync"}},
        "gpu_launches": engine.launches_per_step * R * args.steps,
        "clocks": clocks, "roofline": roofline,
    }
    return line


def run_train(args, rank, world, local_rank, steps=None, warmup=None, workload=None):
    """Training step (forward with targets, sequence losses, backward through the fused operators' backward
    kernels, DDP gradient all-reduce over NCCL, AdamW step), timed like the inference workloads."""
    import torch.distributed as dist
    from camliflow_b200 import native, trainer

    workload = workload or args.workload
    steps, warmup = steps or args.steps, warmup or args.warmup
    H, W, N, iters, B = WORKLOADS[workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    precision = args.train_precision
    from camliflow_b200 import tc
    tc.TRAIN_DENSE = args.train_dense
    torch.backends.cudnn.allow_tf32 = precision != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = precision != "fp32"
    model = build_model(workload).to(dev).train()
    inputs = synthetic_inputs(B, H, W, N, seed=shard_seed(rank))
    g = torch.Generator().manual_seed(1000 + shard_seed(rank))
    inputs["flow_2d"] = torch.randn(B, 2, H, W, generator=g) * 5.0           # SURVEY 8(d) targets
    inputs["flow_3d"] = torch.randn(B, 3, N, generator=g) * 0.1
    pinned = {k: v.pin_memory() for k, v in inputs.items()}
    dev_in = {k: v.to(dev) for k, v in inputs.items()}
    amp = torch.bfloat16 if precision == "bf16" else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        barrier()
        return max_over_ranks(s.elapsed_time(e), world, dev)

    losses = []
    c0 = native.launch_count()
    if args.no_graph:
        ddp = trainer.wrap_ddp(model, dev)
        opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=1e-6)
        mode = "eager: DistributedDataParallel (bucketed all-reduce overlapped with the backward)"

        def step_resident():
            losses.append(trainer.train_step(ddp, opt, dev_in, max_grad_norm=1.0, autocast_dtype=amp))

        def step_e2e():
            batch = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            losses.append(float(trainer.train_step(ddp, opt, batch, max_grad_norm=1.0, autocast_dtype=amp)))   # D2H read of the loss
        step_resident()
        launches = native.launch_count() - c0
    else:
        captured = trainer.CapturedTrainStep(model, dev_in, lr=1e-4, weight_decay=1e-6, max_grad_norm=1.0, autocast_dtype=amp)
        launches = (native.launch_count() - c0) // 4          # 3 warm-up steps + the captured one
        mode = "whole step in ONE CUDA graph: fwd + losses + bwd + one flat gradient all-reduce (NCCL) + clip + AdamW"

        def step_resident():
            losses.append(captured(dev_in))

        def step_e2e():
            losses.append(float(captured(pinned)))            # pinned H2D of the batch, D2H read of the loss

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_resident, steps, warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, steps, warmup)
    pairs = B * world
    grad_bytes = sum(p.numel() for p in model.parameters() if p.requires_grad) * 4
    return {
        "metric": "CamLiRAFT training frame-pairs/sec 960x540+8192pts", "value": pairs * steps / (ms_dev / 1e3),
        "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_dev / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": ("bf16 autocast; dense layers, fused point / correlation operators and losses are fp32 islands "
                           "(hand-written 3xTF32 tensor-core kernels, forward and backward)" if args.train_dense == "tcgen05" else
                           "bf16 autocast (fp32 islands as in the reference: fused point / correlation operators, losses)"),
                  "tf32": "f32 (tf32 library layers)", "fp32": "f32"}[precision], "data": "synthetic",
        "config": {"workload": workload_name(workload), "pairs_per_step": pairs,
                   "parallelism": "dp%d (gradient all-reduce over NCCL, %.1f MB fp32 per step)" % (world, grad_bytes / 1e6),
                   "l2": "working set (activations of %d iterations) larger than L2" % iters,
                   "step": mode, "precision": precision, "grad_clip": 1.0, "final_loss": float(losses[-1]),
                   "dense_layers": ("tcgen05: grad.DenseFn -- camli_conv_gemm forward, camli_transpose_split + camli_conv_gemm (mirrored "
                                    "weights) + camli_conv_wgrad backward, fp32-accurate; layers outside the kernels' coverage "
                                    "(stride 2, C_in % 4 != 0) through cuDNN" if args.train_dense == "tcgen05" else
                                    "library: cuDNN / cuBLAS under autograd (bf16 under autocast)")},
        "e2e": {"value": pairs * steps / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": sum(v.numel() * 4 for v in pinned.values()), "d2h_bytes_per_step": 4},
        "gpu_launches": launches * steps, "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs-per-step", type=int, default=8,
                    help="forwards (of the workload's batch) streamed through the engines per step")
    ap.add_argument("--conv-precision", default="fp32", choices=["fp32", "tf32"],
                    help="library (cuDNN) convolution precision of the inference workloads.  fp32: strict fp32, the mode the "
                         "EPE parity tests run in; tf32: torch's default (fails the EPE bar, kept for A/B only)")
    ap.add_argument("--train-precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="training workloads: bf16 autocast (BASELINE config 5), tf32 library layers (torch default, what the "
                         "reference trains with when amp is off) or strict fp32")
    ap.add_argument("--concurrent", type=int, default=3,
                    help="engines (CUDA graphs of batch B) in flight for the throughput figures (EnginePool); 1 = single engine")
    ap.add_argument("--train-dense", default="tcgen05", choices=["tcgen05", "library"],
                    help="dense layers of the training step: hand-written tensor-core kernels forward + backward (default), or "
                         "cuDNN / cuBLAS under autograd (A/B)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-training-block", action="store_true",
                    help="skip the short BASELINE-config-5 training measurement appended to the default (c2) line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 8))
        base = run_cpu_reference(args.workload, steps=steps, warmup=min(args.warmup, 1))
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "no CPU restatement of CamLiPWC (workload c3) is kept"}))
            return
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "pairs_per_step": WORKLOADS[args.workload][4],
                       "note": "CPU path of the reference algorithm (oracle port), a step = one forward"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload in TRAIN_WORKLOADS:
        line = run_train(args, rank, world, local_rank)
    else:
        line = run_ours(args, rank, world, local_rank)
        if args.workload == "c2" and not args.no_training_block:
            # BASELINE config 5 in the same run: the gradient all-reduce is the only collective of this project
            torch.cuda.empty_cache()
            tr = run_train(args, rank, world, local_rank, steps=4, warmup=3, workload="c5")
            line["training"] = {k: tr[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config",
                                                   "e2e", "gpu_launches")}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = run_cpu_reference(args.workload, steps=1, warmup=1)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
