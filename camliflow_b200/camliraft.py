"""CamLiRAFT model wrapper (reference models/camliraft.py:32-73): pad to a multiple of 8,
ImageNet normalisation, inverse-depth scaling of the clouds, the fused core, and the way back.
The sequence losses of models/losses.py and the metric bookkeeping of models/base.py (device-side accumulation,
one all-reduce at read-out: camliflow_b200/base.py) are evaluated when targets are supplied."""
import torch
import torch.nn as nn

from .base import FlowModel
from .camliraft_core import CamLiRAFT_Core
from .ids import paral2persp, persp2paral
from .losses import calc_sequence_loss_2d, calc_sequence_loss_3d
from .utils import InputPadder


class CamLiRAFT(FlowModel):
    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.core = CamLiRAFT_Core(cfgs)
        self.channels_last = False      # set by FlowEngine: run the 2-D branch in NHWC (cuDNN's native layout)
        self.register_buffer("_mean", torch.tensor([123.675, 116.280, 103.530]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("_std", torch.tensor([58.395, 57.120, 57.375]).view(1, 3, 1, 1), persistent=False)

    def train(self, mode=True):
        super().train(mode)
        if self.cfgs.freeze_bn:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self

    def predictions(self, inputs):
        """Lists of per-iteration predictions: ([B,2,H,W]...], [[B,3,N]...])."""
        images = inputs["images"].float()
        pc1, pc2 = inputs["pcs"][:, :3].float(), inputs["pcs"][:, 3:].float()
        intr = inputs["intrinsics"].float()
        padder = InputPadder(images.shape, x=8)
        image1, image2 = padder.pad(images[:, :3], images[:, 3:])
        image1 = (image1 - self._mean) / self._std
        image2 = (image2 - self._mean) / self._std
        if self.channels_last:
            image1 = image1.contiguous(memory_format=torch.channels_last)
            image2 = image2.contiguous(memory_format=torch.channels_last)
        H, W = image1.shape[-2:]
        persp = {"projection_mode": "perspective", "sensor_h": H, "sensor_w": W,
                 "f": intr[:, 0], "cx": intr[:, 1], "cy": intr[:, 2]}
        qh, qw = round(H / 32), round(W / 32)
        paral = {"projection_mode": "parallel", "sensor_h": qh, "sensor_w": qw,
                 "cx": (qw - 1) / 2, "cy": (qh - 1) / 2}
        pc1 = persp2paral(pc1, persp, paral)
        pc2 = persp2paral(pc2, persp, paral)
        preds_2d, preds_3d = self.core(image1, image2, pc1, pc2, paral)
        preds_2d = [padder.unpad(p) for p in preds_2d]
        base = paral2persp(pc1, persp, paral)
        preds_3d = [paral2persp(pc1 + p, persp, paral) - base for p in preds_3d]
        return preds_2d, preds_3d

    def forward(self, inputs):
        """Inference: the final flows.  With targets in `inputs` (`flow_2d` [B,2|3,H,W], `flow_3d` [B,3|4,N]) also
        the training loss of the reference (models/camliraft.py:80-86) in `self.loss`, `self.loss2d`, `self.loss3d`."""
        preds_2d, preds_3d = self.predictions(inputs)
        if "flow_2d" in inputs and "flow_3d" in inputs:
            self.loss2d = calc_sequence_loss_2d(preds_2d, inputs["flow_2d"].float(), self.cfgs.loss2d)
            self.loss3d = calc_sequence_loss_3d(preds_3d, inputs["flow_3d"].float(), self.cfgs.loss3d)
            self.loss = self.loss2d + self.loss3d
            self.update_metrics("loss", self.loss)
            self.update_metrics("loss2d", self.loss2d)
            self.update_metrics("loss3d", self.loss3d)
            self.update_2d_metrics(preds_2d[-1], inputs["flow_2d"].float())
            self.update_3d_metrics(preds_3d[-1], inputs["flow_3d"].float())
            if "occ_mask_3d" in inputs:
                self.update_3d_metrics(preds_3d[-1], inputs["flow_3d"].float(), inputs["occ_mask_3d"])
        return {"flow_2d": preds_2d[-1], "flow_3d": preds_3d[-1]}

    @staticmethod
    def is_better(curr_metrics, best_metrics):
        return best_metrics is None or curr_metrics["epe2d"] < best_metrics["epe2d"]
