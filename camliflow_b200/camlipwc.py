"""CamLiPWC model wrapper (reference models/camlipwc.py:31-106): /255, resize to a multiple of 64,
inverse-depth scaling, FPS pyramid, encode both frames, fused coarse-to-fine decode, and back."""
import torch.nn as nn

from .base import FlowModel
from .camlipwc_core import CamLiPWC_Core
from .ids import paral2persp, persp2paral
from .losses import calc_pyramid_loss_2d, calc_pyramid_loss_3d
from .utils import build_pc_pyramid, resize_flow2d, resize_to_64x


class CamLiPWC(FlowModel):
    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.core = CamLiPWC_Core(cfgs.pwc2d, cfgs.pwc3d, cfgs.fusion)

    def train(self, mode=True):
        super().train(mode)
        if self.cfgs.freeze_bn:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self

    def predictions(self, inputs):
        """Per-level predictions, finest first: ([B,2,H',W'] x5, [B,3,N_l] x5), H' x W' = the 64-aligned size."""
        images = inputs["images"].float() / 255.0
        pc1, pc2 = inputs["pcs"][:, :3].float(), inputs["pcs"][:, 3:].float()
        intr = inputs["intrinsics"].float()
        origin_h, origin_w = images.shape[2:]
        images = resize_to_64x(images, None)[0]
        image1, image2 = images[:, :3], images[:, 3:]
        persp = {"projection_mode": "perspective", "sensor_h": origin_h, "sensor_w": origin_w,
                 "f": intr[:, 0], "cx": intr[:, 1], "cy": intr[:, 2]}
        qh, qw = round(image1.shape[-2] / 32), round(image1.shape[-1] / 32)
        paral = {"projection_mode": "parallel", "sensor_h": qh, "sensor_w": qw, "cx": (qw - 1) / 2, "cy": (qh - 1) / 2}
        pc1 = persp2paral(pc1, persp, paral)
        pc2 = persp2paral(pc2, persp, paral)
        xyzs1, xyzs2, sample_indices1, _ = build_pc_pyramid(pc1, pc2, [4096, 2048, 1024, 512, 256])
        feats1_2d, feats1_3d = self.core.encode(image1, xyzs1)
        feats2_2d, feats2_3d = self.core.encode(image2, xyzs2)
        flows_2d, flows_3d = self.core.decode(xyzs1, xyzs2, feats1_2d, feats2_2d, feats1_3d, feats2_3d, paral)
        flows_3d = [paral2persp(xyz1 + f, persp, paral) - paral2persp(xyz1, persp, paral)
                    for xyz1, f in zip(xyzs1, flows_3d)]
        return flows_2d, flows_3d, (origin_h, origin_w), sample_indices1

    def forward(self, inputs):
        """Inference: the finest-level flows.  With the targets `flow_2d` / `flow_3d` in `inputs` also the reference's
        pyramid training loss in `self.loss` and its metric bookkeeping (models/camlipwc.py:82-100)."""
        flows_2d, flows_3d, (h, w), sample_indices1 = self.predictions(inputs)
        final_2d, final_3d = resize_flow2d(flows_2d[0], h, w), flows_3d[0]
        if "flow_2d" in inputs and "flow_3d" in inputs:
            target_2d, target_3d = inputs["flow_2d"].float(), inputs["flow_3d"].float()
            self.loss2d = calc_pyramid_loss_2d(flows_2d, target_2d, self.cfgs.loss2d)
            self.loss3d = calc_pyramid_loss_3d(flows_3d, target_3d, self.cfgs.loss3d, sample_indices1)
            self.loss = self.loss2d + self.loss3d
            self.update_metrics("loss", self.loss)
            self.update_metrics("loss2d", self.loss2d)
            self.update_metrics("loss3d", self.loss3d)
            self.update_2d_metrics(final_2d, target_2d)
            self.update_3d_metrics(final_3d, target_3d)
            if "occ_mask_3d" in inputs:
                self.update_3d_metrics(final_3d, target_3d, inputs["occ_mask_3d"])
        return {"flow_2d": final_2d, "flow_3d": final_3d}

    @staticmethod
    def is_better(curr_metrics, best_metrics):
        return best_metrics is None or curr_metrics["epe2d"] < best_metrics["epe2d"]
