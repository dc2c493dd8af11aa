"""Sequence losses of CamLiRAFT (reference models/losses.py:64-119): gamma-weighted sum over the
per-iteration predictions of the masked l2-norm / l1 / robust end-point error."""
import torch


def _sequence_loss(preds, target, n_flow, cfgs):
    mask = target[:, n_flow] > 0 if target.shape[1] == n_flow + 1 else torch.ones_like(target[:, 0], dtype=torch.bool)
    total = 0
    for i, pred in enumerate(preds):
        diff = pred - target[:, :n_flow]
        if cfgs.order == "l2-norm":
            err = torch.linalg.norm(diff, dim=1)
        elif cfgs.order == "l1":
            err = diff.abs().sum(dim=1)
        elif cfgs.order == "robust":
            err = torch.pow(diff.abs().sum(dim=1) + 0.01, 0.4)
        else:
            raise ValueError(cfgs.order)
        total = total + cfgs.gamma ** (len(preds) - i - 1) * err[mask].mean()
    return total


def calc_sequence_loss_2d(flow_preds, target, cfgs):
    """flow_preds: list of [B,2,H,W]; target [B,2(+1 valid mask),H,W] (losses.py:64-90)."""
    return _sequence_loss(flow_preds, target, 2, cfgs)


def calc_sequence_loss_3d(flow_preds, target, cfgs):
    """flow_preds: list of [B,3,N]; target [B,3(+1 valid mask),N] (losses.py:93-119)."""
    return _sequence_loss(flow_preds, target, 3, cfgs)
