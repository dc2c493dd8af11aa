#!/usr/bin/env python
"""Benchmark of the CamLiRAFT hot path (BASELINE.json metric: frame-pairs/s, 960x540 RGB +
8192 points, 12 GRU iterations, batch 1 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N>1); the batch of frame pairs is sharded over the ranks
with no data-path collective ("scaling": "weak").  Prints ONE JSON line on rank 0.

* value        : pairs/s with the inputs resident in HBM (device-timed, CUDA events, max over ranks)
* e2e          : the same end to end from HOST tensors (pinned H2D + forward + D2H for every pair) through the public
                 throughput API: EnginePool.pipelined (3 batch-1 CUDA graphs in flight over shared weights); the
                 single-engine pipelined and the blocking call-per-pair figures are reported beside it
* roofline     : the dominant hand-written kernel, timed per launch with CUDA events on its stream
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
    # name: (H, W, N points, GRU iterations, pairs per GPU)
    "c2": (540, 960, 8192, 12, 1),
    "small": (160, 224, 8192, 3, 1),
    # BASELINE config 4 (`--workload c4`): 32 iterations, 32 pairs over 8 GPUs = 4 pairs per GPU
    "c4": (540, 960, 8192, 32, 4),
    # BASELINE config 5 (secondary; `--workload c5`): training step, n_iters_train = 10, 2 pairs per GPU
    "c5": (540, 960, 8192, 10, 2),
    "c5small": (160, 224, 8192, 3, 1),
}
TRAIN_WORKLOADS = ("c5", "c5small")
METRIC = "CamLiRAFT frame-pairs/sec 960x540+8192pts"


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
    """Times the reference algorithm's CPU path: oracle port, torch fallbacks for FPS / k-NN
    (what models/csrc/wrapper.py does for CPU tensors), all host threads, eval mode, every
    iteration's prediction materialised like the reference."""
    from oracle import camliraft_oracle as co
    H, W, N, iters, B = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = co.make_params(co.param_spec("camliraft"), seed=0)
    inp = co.synthetic_inputs(B, H, W, N, seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        co.camliraft_forward(P, inp["images"], inp["pcs"], inp["intrinsics"], n_iters=iters, index_impl="fallback",
                             all_iters=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": B / sec, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d full forward(s) of %d pair(s), %dx%d + %d pts, %d iters, after %d warm-up"
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


def pairs_per_step(B, world):
    return B * world


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from camliflow_b200 import native, ops
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.engine import FlowEngine
    from camliflow_b200.init import seed_module_

    H, W, N, iters, B = WORKLOADS[args.workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    strict = args.conv_precision == "fp32"
    torch.backends.cudnn.allow_tf32 = not strict
    torch.backends.cuda.matmul.allow_tf32 = False      # GEMMs (1x1 layers, Linear) are always fp32
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=iters)), seed=0)
    engine = FlowEngine(model, B, H, W, N, device=dev, use_graph=not args.no_graph)
    inputs = synthetic_inputs(B, H, W, N, seed=shard_seed(rank))   # per-rank shard of the batch of pairs
    pinned = {k: v.pin_memory() for k, v in inputs.items()}
    flush = torch.zeros(192 * 1024 * 1024 // 4, device=dev)   # 192 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        with torch.cuda.stream(engine.stream):
            for _ in range(steps):
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
    ms_dev = timed(engine.step, args.steps, args.warmup)
    if os.environ.get("CAMLI_PROFILER_RANGE"):
        torch.cuda.profiler.stop()
    ms_e2e = timed(lambda: engine(pinned), args.steps, args.warmup)

    # the same end-to-end call in throughput mode: H2D of pair i+1 and D2H of pair i-1 overlap the compute of pair i
    # (every pair still pays both transfers; wall clock on the host around K pairs, barrier + synchronize both sides)
    def piped(n):
        for out in engine.pipelined(pinned for _ in range(n)):
            pass
    piped(args.warmup)
    barrier()
    t0 = time.perf_counter()
    piped(args.steps)
    barrier()
    ms_pipe = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)

    # serving-style concurrency: two engines (two CUDA graphs over the same weights) fed round-robin, so the
    # latency-bound phases of one pair overlap another pair's work; every forward is still batch B
    conc = None
    if args.concurrent > 1:
        from camliflow_b200.engine import EnginePool
        pool = EnginePool(model, args.concurrent, B, H, W, N, device=dev, use_graph=not args.no_graph)

        def pooled(n):
            for out in pool.pipelined(pinned for _ in range(n)):
                pass
        pooled(args.warmup + args.concurrent)
        barrier()
        t0 = time.perf_counter()
        pooled(args.steps)
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
                pool.engines[i % args.concurrent].step()
            for eng in pool.engines:
                ctl.wait_stream(eng.stream)
            if timed_pair is not None:
                timed_pair[1].record(ctl)
        resident(args.warmup)
        barrier()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        resident(args.steps, ev)
        barrier()
        ms_pool_dev = max_over_ranks(ev[0].elapsed_time(ev[1]), world, dev)
        conc = {"value": pairs_per_step(B, world) * args.steps / (ms_pool / 1e3), "unit": "pairs/s", "ms_per_step": ms_pool / args.steps,
                "engines": args.concurrent,
                "mode": "EnginePool.pipelined: %d CUDA graphs of batch %d in flight, round-robin; each pair pays its own pinned "
                        "H2D + D2H (host wall clock)" % (args.concurrent, B)}

    clocks = sampler.stop() if rank == 0 else None      # sampled over every timed region above

    # per-launch timing of the dominant hand-written kernel (eager pass, events on the launch stream)
    roofline = None
    if rank == 0:
        ops.profile_begin()
        with torch.cuda.stream(engine.stream), torch.no_grad():
            for _ in range(2):
                l2_flush(flush)
                engine._forward_static()
        engine.stream.synchronize()
        roofline = ops.profile_end(peaks_path=os.path.join(ROOT, "MEASURED_PEAKS.json"))

    h2d, d2h = engine.io_bytes()
    pairs = B * world
    line = {
        "metric": METRIC, "value": pairs * args.steps / (ms_dev / 1e3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: CamLiRAFT fusion %dx%d RGB + %d pts, %d iters, batch %d per GPU"
                               % (args.workload, W, H, N, iters, B),
                   "pairs_per_step": pairs, "cuda_graph": engine.graph is not None,
                   "l2": "192 MiB flush write before every timed step of `value` (outside the event pair); e2e: inputs re-copied "
                         "every step and a per-step working set (355 MB volume pyramid + activations) larger than L2",
                   "conv_precision": ("fp32 (cudnn.allow_tf32=False: the mode the parity tests run in)" if strict else
                                      "cuDNN default (TF32 allowed, as torch default in the reference)"),
                   "intermediate_predictions": False},
        # headline: the public end-to-end call in throughput mode (FlowEngine.pipelined): every pair pays its pinned
        # H2D and its D2H inside the timed region, overlapped with the compute of the neighbouring pairs; host wall
        # clock.  `synchronous` = one blocking FlowEngine.__call__ per pair (latency mode), CUDA-event timed.
        "e2e": {"value": pairs * args.steps / (ms_pipe / 1e3), "unit": "pairs/s", "ms_per_step": ms_pipe / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "FlowEngine.pipelined: pinned H2D of pair i+1 and D2H of pair i-1 overlap the forward of pair i",
                "synchronous": {"value": pairs * args.steps / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                                "mode": "FlowEngine.__call__ per pair: pinned H2D -> forward -> D2H -> host sync"}},
        "gpu_launches": engine.launches_per_step * args.steps,
        "clocks": clocks, "roofline": roofline,
    }
    if conc is not None:
        # `value`: the same pool with resident inputs, CUDA-event timed; the single-engine (one pair at a time) figure,
        # whose ms_per_step is the latency of a pair, stays beside it
        line["single_engine"] = {"value": line["value"], "unit": "pairs/s", "ms_per_step": line["ms_per_step"],
                                 "note": "one CUDA graph at a time: ms_per_step is the latency of a frame pair"}
        line["value"] = pairs * args.steps / (ms_pool_dev / 1e3)
        line["ms_per_step"] = ms_pool_dev / args.steps
        line["config"]["engines_in_flight"] = args.concurrent
        line["config"]["l2"] = ("value / e2e: %d CUDA graphs in flight, per-engine working set 355 MB volume pyramid + activations "
                                "(larger than L2), inputs re-copied every step for e2e; single_engine: 192 MiB flush write "
                                "before every timed step (outside the event pair)" % args.concurrent)
        # headline end-to-end throughput: the serving-style pool (every forward still a batch-B graph, every pair pays its
        # own transfers); the single-engine pipelined and synchronous figures stay beside it
        single = {k: line["e2e"][k] for k in ("value", "unit", "ms_per_step", "mode")}
        line["e2e"].update({"value": conc["value"], "ms_per_step": conc["ms_per_step"], "mode": conc["mode"],
                            "engines_in_flight": conc["engines"], "single_engine_pipelined": single})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu_reference(args.workload, steps=1, warmup=1)
    return line


def run_train(args, rank, world, local_rank):
    """Training step (forward with targets, sequence losses, backward through the fused operators' backward
    kernels, DDP gradient all-reduce over NCCL, AdamW step), timed like the inference workloads."""
    import torch.distributed as dist
    from camliflow_b200 import native, trainer
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_

    H, W, N, iters, B = WORKLOADS[args.workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    strict = args.conv_precision == "fp32"
    torch.backends.cudnn.allow_tf32 = not strict
    torch.backends.cuda.matmul.allow_tf32 = not strict
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=iters)), seed=0).to(dev).train()
    ddp = trainer.wrap_ddp(model, dev)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=1e-6)
    inputs = synthetic_inputs(B, H, W, N, seed=shard_seed(rank))
    g = torch.Generator().manual_seed(1000 + shard_seed(rank))
    inputs["flow_2d"] = torch.randn(B, 2, H, W, generator=g) * 5.0           # SURVEY 8(d) targets
    inputs["flow_3d"] = torch.randn(B, 3, N, generator=g) * 0.1
    pinned = {k: v.pin_memory() for k, v in inputs.items()}
    dev_in = {k: v.to(dev) for k, v in inputs.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        return max_over_ranks(s.elapsed_time(e), world, dev)

    losses = []

    def step_resident():
        losses.append(trainer.train_step(ddp, opt, dev_in))

    def step_e2e():
        batch = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        losses.append(float(trainer.train_step(ddp, opt, batch)))            # D2H read of the loss

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    c0 = native.launch_count()
    ms_dev = timed(step_resident, args.steps, args.warmup)
    launches = (native.launch_count() - c0) // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    pairs = B * world
    return {
        "metric": "CamLiRAFT training frame-pairs/sec 960x540+8192pts", "value": pairs * args.steps / (ms_dev / 1e3),
        "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: CamLiRAFT training step (fwd + sequence losses + bwd + AdamW) %dx%d RGB + %d pts, "
                               "%d iters, batch %d per GPU" % (args.workload, W, H, N, iters, B),
                   "pairs_per_step": pairs, "parallelism": "dp%d (DDP gradient all-reduce over NCCL)" % world,
                   "l2": "working set (activations of %d iterations) larger than L2" % iters,
                   "conv_precision": ("fp32 (strict)" if strict else "tf32 for the cuDNN / cuBLAS layers of the autograd path (torch's "
                                      "default, what the reference trains with); the fused operators stay fp32"),
                   "final_loss": float(losses[-1])},
        "e2e": {"value": pairs * args.steps / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": sum(v.numel() * 4 for v in pinned.values()), "d2h_bytes_per_step": 4},
        "gpu_launches": launches * args.steps, "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--conv-precision", default=None, choices=["fp32", "tf32"],
                    help="library (cuDNN / cuBLAS) convolution precision.  fp32: strict fp32, the mode the EPE parity tests "
                         "run in (default for the inference workloads); tf32: torch's default (allow_tf32=True), what the "
                         "reference trains with out of the box (default for the training workloads c5*)")
    ap.add_argument("--concurrent", type=int, default=3,
                    help="engines (CUDA graphs of batch B) in flight for the end-to-end throughput figure e2e.value "
                         "(EnginePool.pipelined); 1 = single engine (FlowEngine.pipelined)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.conv_precision is None:
        args.conv_precision = "tf32" if args.workload in TRAIN_WORKLOADS else "fp32"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        H, W, N, iters, B = WORKLOADS[args.workload]
        steps = max(1, min(args.steps, 8))
        base = run_cpu_reference(args.workload, steps=steps, warmup=min(args.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: CamLiRAFT fusion %dx%d RGB + %d pts, %d iters, batch %d (CPU path)"
                                   % (args.workload, W, H, N, iters, B)},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = (run_train if args.workload in TRAIN_WORKLOADS else run_ours)(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
