"""Training step of the fused path (BASELINE config 5; reference train.py:143-171 minus its logging):
forward with targets -> losses -> backward -> gradient all-reduce -> clip -> optimizer step.

Data parallel over frame pairs: one process per GPU, each rank owns its pairs.  The ONLY data-path collective is the
all-reduce of the 8.4 M fp32 gradients (33.5 MB) over NCCL (NVLink 5 / NVSwitch) -- P2P on, unlike the reference's
environment overrides (train.py:30-31).  Two ways to run it:

* `wrap_ddp` + `train_step`: torch DistributedDataParallel, bucketed all-reduce overlapped with the backward, as in the
  reference (train.py:101).
* `CapturedTrainStep`: the whole step -- forward, losses, backward, ONE flat all-reduce of a contiguous gradient
  buffer, clipping, AdamW -- captured in a single CUDA graph.  A CamLiRAFT step is ~13 k kernel launches; eagerly the
  host needs ~250 ms to issue what the GPU executes in ~130 ms (scripts/trace_train.py), so the graph removes the
  bound.  At 33.5 MB the un-overlapped all-reduce costs ~0.1 ms on NVSwitch against a >100 ms step.
"""
import os

import torch
import torch.distributed as dist


def wrap_ddp(model, device=None, bucket_cap_mb=16):
    """DistributedDataParallel around `model` when a process group is up (identity for one process).  The
    fused operators are ordinary autograd Functions, so DDP's bucket hooks fire as their gradients land."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP
    ids = None if device is None or torch.device(device).type != "cuda" else [torch.device(device).index]
    return DDP(model, device_ids=ids, bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)


def unwrap(model):
    return model.module if hasattr(model, "module") else model


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def train_step(model, optimizer, inputs, max_grad_norm=None, autocast_dtype=None):
    """One optimisation step; returns the (detached) loss tensor.  `inputs` must carry the targets
    `flow_2d` / `flow_3d` (models/camliraft.py:80-86).  `autocast_dtype` (torch.bfloat16): forward and loss under
    autocast as in train.py:147-149 -- dense layers the kernels do not cover run in bf16 (cuDNN / cuBLAS), the ones they do
    (grad.DenseFn), the fused point / correlation operators and the losses stay fp32 (camliflow_b200/grad.py:f32); bf16 needs no GradScaler."""
    from . import grad
    grad.clear_dense_cache()                 # split weights of the dense-layer nodes live for one step
    optimizer.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
        model(inputs)
        loss = unwrap(model).loss
    loss.backward()
    if max_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    return loss.detach()


def flatten_gradients(model):
    """Gives every trainable parameter a `.grad` that is a view into ONE contiguous fp32 buffer (returned), so the
    data-parallel reduction is a single NCCL call and zeroing the gradients is a single memset."""
    params = [p for p in model.parameters() if p.requires_grad]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    offset = 0
    for p in params:
        p.grad = flat[offset:offset + p.numel()].view_as(p)
        offset += p.numel()
    return flat


def allreduce_mean_(flat):
    """In-place mean over the ranks of the flat gradient buffer (no-op for one process)."""
    if _world() > 1:
        dist.all_reduce(flat)
        flat.div_(_world())
    return flat


class CapturedTrainStep:
    """The whole training step as one CUDA graph over static input buffers.

        step = CapturedTrainStep(model, example_inputs, lr=..., autocast_dtype=torch.bfloat16)
        loss = step(inputs)        # copies `inputs` into the static buffers, replays, returns the device loss tensor

    Model must be in train mode on its device, NOT wrapped in DDP (the all-reduce is explicit, see module docstring).
    The optimizer is AdamW(capturable=True).  `use_graph=False` runs the same step eagerly (A/B, debugging)."""

    def __init__(self, model, example_inputs, lr=1e-4, weight_decay=1e-6, max_grad_norm=1.0, autocast_dtype=None,
                 use_graph=True, warmup=3):
        self.model, self.max_grad_norm, self.autocast_dtype = model, max_grad_norm, autocast_dtype
        dev = next(model.parameters()).device
        self.static_in = {k: v.to(dev).clone() for k, v in example_inputs.items()}
        self.flat = flatten_gradients(model)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, capturable=use_graph, foreach=True)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.graph = None
        model.track_metrics = False
        if not use_graph:
            return
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # allocator / cuDNN autotune / NCCL warm-up outside the capture
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        from . import grad
        grad.clear_dense_cache()                          # (its entries live in the graph's memory pool)

    def _step(self):
        from . import grad
        grad.clear_dense_cache()                          # every weight is split once per step, inside the step
        self.flat.zero_()
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            self.model(self.static_in)
            loss = self.model.loss
        # (in-place weight-gradient accumulation of the 1x1 layers: measured, 97.5 -> 101.3 ms per step, so off by default)
        with grad.fused_wgrad_accumulation(os.environ.get("CAMLI_FUSE_WGRAD", "0") == "1"):
            loss.backward()
        allreduce_mean_(self.flat)
        if self.max_grad_norm is not None:
            # clip_grad_norm_ on the flat buffer: one norm, one scale, no host round trip
            norm = torch.linalg.vector_norm(self.flat)
            self.flat.mul_(torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0))
        self.opt.step()
        self.loss.copy_(loss.detach())

    def __call__(self, inputs):
        for k, dst in self.static_in.items():
            dst.copy_(inputs[k], non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.loss
