"""Loss / metric bookkeeping of the model wrappers (reference models/base.py:6-94).

Same surface (`update_metrics`, `update_2d_metrics`, `update_3d_metrics`, `get_metrics`, `clear_metrics`,
`get_loss`) and the same definitions (end-point error, 1 px / 5 cm accuracy, 3 px & 5 % outliers), scheduled
differently: the reference pulls every metric to the host (`.item()`) and all-reduces value and count
separately -- about twenty blocking 4-byte all-reduces per training step (models/base.py:16-32,
models/utils.py:272-278).  Here every update is a handful of device-side reductions into one [sum, count] row
per metric; nothing synchronises until `get_metrics()`, which moves the whole table with ONE all-reduce and one
D2H copy."""
import torch
import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.loss = None
        self.track_metrics = True       # False: forward() skips the bookkeeping (whole-step CUDA-graph capture)
        self._metric_rows = {}          # name -> device tensor [2] = (sum, count), float64

    def clear_metrics(self):
        self._metric_rows = {}

    @torch.no_grad()
    def update_metrics(self, name, var, mask=None):
        """Accumulates sum(var[mask]) and its element count under `name` (models/base.py:16-32).  `var`: tensor
        (any shape; bool is counted as 0/1) or a python number (count 1)."""
        if not self.track_metrics:
            return
        if not isinstance(var, torch.Tensor):
            var = torch.tensor(float(var))
        v = var.detach().to(torch.float64)
        if mask is not None:
            m = mask.to(torch.float64)
            row = torch.stack([(v * m).sum(), m.sum()])
        else:
            row = torch.stack([v.sum(), torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device)])
        prev = self._metric_rows.get(name)
        self._metric_rows[name] = row if prev is None else prev + row.to(prev.device)

    def get_metrics(self):
        """{name: sum / count} over every rank (one flat all-reduce); metrics with no valid element are omitted,
        as in the reference (models/base.py:24-25)."""
        if not self._metric_rows:
            return {}
        names = sorted(self._metric_rows)
        dev = next((r.device for r in self._metric_rows.values() if r.is_cuda), torch.device("cpu"))
        table = torch.stack([self._metric_rows[n].to(dev) for n in names])
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(table)
        table = table.cpu()
        return {n: float(table[i, 0] / table[i, 1]) for i, n in enumerate(names) if table[i, 1] > 0}

    def get_loss(self):
        if self.loss is None:
            raise ValueError("Loss is empty.")
        return self.loss

    @staticmethod
    def is_better(curr_metrics, best_metrics):
        raise RuntimeError("Function `is_better` must be implemented.")


class FlowModel(BaseModel):
    """models/base.py:50-94."""

    @torch.no_grad()
    def update_2d_metrics(self, pred, target):
        if not self.track_metrics:
            return
        if target.shape[1] == 3:                         # sparse ground truth: channel 2 is the validity mask
            mask, target = target[:, 2] > 0, target[:, :2]
        else:
            mask = None
        epe = torch.linalg.norm(pred - target, dim=1)
        self.update_metrics("epe2d", epe, mask)
        self.update_metrics("acc2d_1px", epe < 1.0, mask)
        mag = torch.linalg.norm(target, dim=1) + 1e-5
        self.update_metrics("outlier2d", torch.logical_and(epe > 3.0, epe / mag > 0.05), mask)

    @torch.no_grad()
    def update_3d_metrics(self, pred, target, occ_mask=None):
        if not self.track_metrics:
            return
        if target.shape[1] == 4:
            mask, target = target[:, 3] > 0, target[:, :3]
        else:
            mask = None
        epe = torch.linalg.norm(pred - target, dim=1)
        acc = epe < 0.05
        if occ_mask is not None:
            noc = occ_mask == 0
            mask = noc if mask is None else torch.logical_and(noc, mask)
            self.update_metrics("epe3d_noc", epe, mask)
            self.update_metrics("acc3d_5cm_noc", acc, mask)
        else:
            self.update_metrics("epe3d", epe, mask)
            self.update_metrics("acc3d_5cm", acc, mask)
