"""Fixed-shape inference runner: pinned host staging buffers, static device buffers and the
whole forward captured in ONE CUDA graph, so a B=1 frame pair costs one graph launch instead
of ~1.5 k eager launches.  `FlowEngine.__call__` is the public end-to-end
call: host tensors in, host tensors out."""
import os

import torch

from . import native


class FlowEngine:
    def __init__(self, model, batch, height, width, n_points, device="cuda:0", use_graph=True, warmup=2,
                 channels_last=True, tile_policy=None):
        """`tile_policy`: "latency" for an engine that runs ONE graph at a time (64-wide tiles for the layers that would
        otherwise leave half of the SMs idle), "throughput" (default of ops.TILE_POLICY) for engines that share the GPU
        with other graphs in flight (EnginePool)."""
        self.device = torch.device(device)
        self.tile_policy = tile_policy
        self.model = model.to(self.device).eval()
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
            self.model.channels_last = True
        self.shape = (batch, height, width, n_points)
        mk = lambda *s: torch.zeros(*s, dtype=torch.float32)   # noqa: E731
        self.host_in = {"images": mk(batch, 6, height, width).pin_memory(), "pcs": mk(batch, 6, n_points).pin_memory(),
                        "intrinsics": mk(batch, 3).pin_memory()}
        self.dev_in = {k: v.to(self.device) for k, v in self.host_in.items()}
        self.host_out = {"flow_2d": mk(batch, 2, height, width).pin_memory(), "flow_3d": mk(batch, 3, n_points).pin_memory()}
        self.dev_out = None
        self.graph = None
        self.launches_per_step = 0
        # (a high-priority main stream was measured: it delays the FPS on the side stream and lengthens the pre-loop phase)
        self.stream = torch.cuda.Stream(self.device, priority=int(os.environ.get("CAMLI_MAIN_PRIORITY", "0")))
        self._prepare(use_graph, warmup)

    # ------------------------------------------------------------------ setup
    def _forward_static(self):
        from . import ops
        old = ops.TILE_POLICY
        if self.tile_policy is not None:
            ops.TILE_POLICY = self.tile_policy
        try:
            with torch.no_grad():
                out = self.model(self.dev_in)
        finally:
            ops.TILE_POLICY = old
        if self.dev_out is None:
            self.dev_out = {k: torch.empty_like(v) for k, v in out.items()}
        for k in self.dev_out:
            self.dev_out[k].copy_(out[k])

    def _prepare(self, use_graph, warmup):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            # inputs must be physically plausible during warm-up (log(z) of the IDS transform)
            g = torch.Generator(device=self.device).manual_seed(1234)
            pcs = torch.rand(self.dev_in["pcs"].shape, generator=g, device=self.device)
            pcs[:, 0::3] = (pcs[:, 0::3] - 0.5) * 8.0
            pcs[:, 1::3] = (pcs[:, 1::3] - 0.5) * 4.0
            pcs[:, 2::3] = pcs[:, 2::3] * 30.0 + 5.0
            self.dev_in["pcs"].copy_(pcs)
            self.dev_in["images"].copy_(torch.rand(self.dev_in["images"].shape, generator=g, device=self.device) * 255)
            self.dev_in["intrinsics"][:] = torch.tensor([1050.0, (self.shape[2] - 1) / 2, (self.shape[1] - 1) / 2])
            for _ in range(max(1, warmup)):
                self._forward_static()
            self.stream.synchronize()
            c0 = native.launch_count()
            self._forward_static()
            self.stream.synchronize()
            self.launches_per_step = native.launch_count() - c0
            if use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._forward_static()
                self.graph = g
        torch.cuda.synchronize(self.device)
        self._use_graph, self._warmup = use_graph, warmup
        self._weights_key = self._weights_fingerprint()

    # ------------------------------------------------------------------ weights are part of the captured graph
    def _weights_fingerprint(self):
        """(storage address, version counter) of every parameter and buffer: the captured graph reads the split / folded
        copies made from these values, so any in-place update, load_state_dict or .to() invalidates it."""
        return tuple((t.data_ptr(), t._version) for t in list(self.model.parameters()) + list(self.model.buffers()))

    def check_weights(self):
        """Raise if the model's weights changed since the graph was captured (an optimizer step, load_state_dict, ...):
        replaying would silently keep computing with the old ones.  Called by the public entry points `__call__` and
        `pipelined`; `step()` itself stays check-free (it is the inner loop of the throughput paths)."""
        if self.graph is not None and self._weights_fingerprint() != self._weights_key:
            raise RuntimeError("FlowEngine: the model's weights changed after the CUDA graph was captured; call "
                               "engine.recapture() (the graph holds split / folded copies of the old values)")

    def recapture(self):
        """Drop the captured graph and the cached split weights, and capture again from the model's current weights."""
        from . import ops, tc
        self.graph = None
        ops._TC_WEIGHTS.clear()
        tc._PLAIN.clear()
        self._prepare(self._use_graph, self._warmup)

    # ------------------------------------------------------------------ execution
    def load(self, inputs):
        """Host -> pinned staging -> device (async on the engine stream)."""
        with torch.cuda.stream(self.stream):
            for k, dst in self.dev_in.items():
                src = inputs[k]
                if src.device.type == "cpu" and not src.is_pinned():
                    self.host_in[k].copy_(src)              # pageable host memory: stage through the pinned buffer
                    dst.copy_(self.host_in[k], non_blocking=True)
                else:
                    dst.copy_(src, non_blocking=True)       # device tensor, or pinned already: DMA straight from the caller's
                                                            # buffer (which must stay untouched until the step has run)

    def step(self):
        """One forward over the resident inputs (no host traffic)."""
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward_static()

    def fetch(self):
        with torch.cuda.stream(self.stream):
            for k, dst in self.host_out.items():
                dst.copy_(self.dev_out[k], non_blocking=True)
        self.stream.synchronize()
        return self.host_out

    def __call__(self, inputs):
        self.check_weights()
        self.load(inputs)
        self.step()
        return self.fetch()

    # ------------------------------------------------------------------ pipelined execution
    def pipelined(self, batches):
        """Throughput mode of the end-to-end call: yields the host outputs of every batch of `batches` (an iterable
        of input dicts of host tensors), with the H2D copy of batch i+1 and the D2H copy of batch i-1 running on a
        copy stream while batch i computes.  Every batch still pays its own pinned H2D / D2H transfers; they are
        overlapped, not skipped.  Double-buffered staging on both sides; the compute stream only ever waits on
        events.  The yielded dict is reused two batches later: consume (or clone) it before advancing twice."""
        dev = self.device
        self.check_weights()
        if not hasattr(self, "_pipe"):
            with torch.cuda.device(dev):
                self._pipe = {
                    "copy": torch.cuda.Stream(dev),
                    "in_host": [{k: torch.empty_like(v).pin_memory() for k, v in self.host_in.items()} for _ in range(2)],
                    "in_dev": [{k: torch.empty_like(v) for k, v in self.dev_in.items()} for _ in range(2)],
                    "out_dev": [{k: torch.empty_like(v) for k, v in self.dev_out.items()} for _ in range(2)],
                    "out_host": [{k: torch.empty_like(v).pin_memory() for k, v in self.host_out.items()} for _ in range(2)],
                }
        P, copy = self._pipe, self._pipe["copy"]
        loaded = [torch.cuda.Event(), torch.cuda.Event()]       # H2D of slot done
        computed = [torch.cuda.Event(), torch.cuda.Event()]     # outputs of slot ready in out_dev
        fetched = [torch.cuda.Event(), torch.cuda.Event()]      # D2H of slot done
        consumed = [torch.cuda.Event(), torch.cuda.Event()]     # compute stream finished reading in_dev[slot]

        def upload(slot, inputs):
            srcs = {}
            for k, dst in P["in_host"][slot].items():
                if inputs[k].is_pinned():
                    srcs[k] = inputs[k]                     # pinned already: no staging memcpy
                else:
                    dst.copy_(inputs[k])
                    srcs[k] = dst
            with torch.cuda.stream(copy):
                copy.wait_event(consumed[slot])
                for k, dst in P["in_dev"][slot].items():
                    dst.copy_(srcs[k], non_blocking=True)
                loaded[slot].record(copy)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        for ev in consumed + fetched:
            ev.record(self.stream)
        upload(0, nxt)
        i = 0
        pending = None                                          # slot whose D2H is in flight
        while nxt is not None:
            slot = i & 1
            nxt = next(it, None)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(loaded[slot])
                for k, dst in self.dev_in.items():
                    dst.copy_(P["in_dev"][slot][k], non_blocking=True)   # device-to-device: microseconds
                consumed[slot].record(self.stream)
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._forward_static()
                self.stream.wait_event(fetched[slot])           # out_dev[slot] no longer being read by the copy stream
                for k, dst in P["out_dev"][slot].items():
                    dst.copy_(self.dev_out[k], non_blocking=True)
                computed[slot].record(self.stream)
            if nxt is not None:
                upload(slot ^ 1, nxt)                           # overlaps with the compute just enqueued
            with torch.cuda.stream(copy):
                copy.wait_event(computed[slot])
                for k, dst in P["out_host"][slot].items():
                    dst.copy_(P["out_dev"][slot][k], non_blocking=True)
                fetched[slot].record(copy)
            if pending is not None:
                fetched[pending].synchronize()
                yield P["out_host"][pending]
            pending = slot
            i += 1
        fetched[pending].synchronize()
        yield P["out_host"][pending]

    def fetch_async(self):
        """Enqueue the D2H copy of the outputs on the engine stream and return the event that marks its end."""
        with torch.cuda.stream(self.stream):
            for k, dst in self.host_out.items():
                dst.copy_(self.dev_out[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def io_bytes(self):
        h2d = sum(v.numel() * v.element_size() for v in self.host_in.values())
        d2h = sum(v.numel() * v.element_size() for v in self.host_out.values())
        return h2d, d2h


class EnginePool:
    """Serving-style throughput runner: `n` FlowEngines over ONE model (shared weights), each with its own
    streams, static buffers and CUDA graph, fed round-robin.  Every forward is still a batch-of-`batch` graph;
    consecutive frame pairs are independent, so the latency-bound phases of one pair (furthest-point sampling:
    16 of 148 SMs for 2 ms; the single-wave kernels of the refinement loop) are filled with another pair's work.
    Each pair pays its own pinned H2D and D2H on its engine's stream; results are yielded in submission order."""

    def __init__(self, model, n, batch, height, width, n_points, device="cuda:0", use_graph=True):
        self.engines = [FlowEngine(model, batch, height, width, n_points, device=device, use_graph=use_graph) for _ in range(n)]

    def pipelined(self, batches):
        n = len(self.engines)
        for eng in self.engines:
            if hasattr(eng, "check_weights"):
                eng.check_weights()
        pending = [None] * n
        for i, inputs in enumerate(batches):
            eng = self.engines[i % n]
            if pending[i % n] is not None:
                pending[i % n].synchronize()
                yield eng.host_out                      # (consume before this engine's next result lands: n batches later)
            eng.load(inputs)
            eng.step()
            pending[i % n] = eng.fetch_async()
            last = i
        if not any(p is not None for p in pending):
            return
        for j in range(last + 1, last + 1 + n):         # drain in submission order
            if pending[j % n] is not None:
                pending[j % n].synchronize()
                pending[j % n] = None
                yield self.engines[j % n].host_out
