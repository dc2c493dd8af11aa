"""Training step of the fused path (BASELINE config 5; reference train.py:150-190 minus its logging):
forward with targets -> sequence losses -> backward -> gradient all-reduce -> optimizer step.

Data parallel over frame pairs: one process per GPU, each rank owns its pairs; the only collective is the
all-reduce of the 8.4 M fp32 gradients (33.5 MB), bucketed and overlapped with the backward by
DistributedDataParallel over NCCL (NVLink / NVSwitch), as in the reference (train.py:79-85) -- but without
its P2P-off environment overrides (train.py:30-31)."""
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


def train_step(model, optimizer, inputs, max_grad_norm=None, autocast_dtype=None):
    """One optimisation step; returns the (detached) loss tensor.  `inputs` must carry the targets
    `flow_2d` / `flow_3d` (models/camliraft.py:80-86).  `autocast_dtype` (torch.bfloat16): forward and loss under
    autocast as in train.py:147-149 -- the dense layers run in bf16, the fused point / correlation operators and
    the losses stay fp32 (camliflow_b200/grad.py:f32); bf16 needs no GradScaler."""
    optimizer.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=autocast_dtype is not None):
        model(inputs)
        loss = unwrap(model).loss
    loss.backward()
    if max_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    return loss.detach()
