"""Sequence losses of CamLiRAFT (reference models/losses.py:64-119): gamma-weighted sum over the
per-iteration predictions of the masked l2-norm / l1 / robust end-point error."""
import torch


def _masked_mean(err, mask):
    """mean of err over mask (None = everywhere) without a data-dependent shape: capturable in a CUDA graph."""
    if mask is None:
        return err.mean()
    m = mask.to(err.dtype)
    return (err * m).sum() / m.sum()


def _sequence_loss(preds, target, n_flow, cfgs):
    mask = target[:, n_flow] > 0 if target.shape[1] == n_flow + 1 else None
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
        total = total + cfgs.gamma ** (len(preds) - i - 1) * _masked_mean(err, mask)
    return total


def calc_sequence_loss_2d(flow_preds, target, cfgs):
    """flow_preds: list of [B,2,H,W]; target [B,2(+1 valid mask),H,W] (losses.py:64-90)."""
    return _sequence_loss(flow_preds, target, 2, cfgs)


def calc_sequence_loss_3d(flow_preds, target, cfgs):
    """flow_preds: list of [B,3,N]; target [B,3(+1 valid mask),N] (losses.py:93-119)."""
    return _sequence_loss(flow_preds, target, 3, cfgs)


def _level_error(diff, order):
    if order == "robust":
        return torch.pow(diff.abs().sum(dim=1) + 0.01, 0.4)
    if order == "l2-norm":
        return torch.linalg.norm(diff, dim=1)
    raise NotImplementedError(order)


def calc_pyramid_loss_2d(flows, target, cfgs):
    """Pyramid loss of CamLiPWC's image branch (losses.py:5-32): every level's prediction [B,2,h,w] is resized
    (bilinear, align_corners, flow scaled) to the target's resolution and compared with it; the per-level means
    are weighted by cfgs.level_weights (finest first)."""
    from .utils import resize_flow2d
    assert len(flows) <= len(cfgs.level_weights)
    mask = target[:, 2] > 0 if target.shape[1] == 3 else None
    total = 0
    for pred, weight in zip(flows, cfgs.level_weights):
        assert pred.shape[1] == 2
        err = _level_error(torch.abs(resize_flow2d(pred, target.shape[2], target.shape[3]) - target[:, :2]), cfgs.order)
        total = total + weight * _masked_mean(err, mask)
    return total


def calc_pyramid_loss_3d(flows, target, cfgs, indices):
    """Pyramid loss of CamLiPWC's point branch (losses.py:35-61): level l's prediction [B,3,N_l] is compared with the
    target gathered at that level's furthest-point sample indices (`indices[l]`, from build_pc_pyramid)."""
    from .utils import batch_indexing
    assert len(flows) <= len(cfgs.level_weights)
    total = 0
    for lvl, (flow, weight) in enumerate(zip(flows, cfgs.level_weights)):
        level_target = batch_indexing(target, indices[lvl])
        err = _level_error(flow - level_target[:, :3], cfgs.order)
        total = total + weight * _masked_mean(err, level_target[:, 3] > 0 if level_target.shape[1] == 4 else None)
    return total
