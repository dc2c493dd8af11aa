"""CamLiRAFT-L, the LiDAR-only model (reference models/camliraft_l.py:7-79; BASELINE config 1): inverse-depth
scaling of the two clouds, the point branch alone (CamLiRAFT_L_Core), and the way back to metric flow."""
import torch.nn as nn

from .camliraft_l_core import CamLiRAFT_L_Core
from .ids import paral2persp, persp2paral
from .losses import calc_sequence_loss_3d


class CamLiRAFT_L(nn.Module):
    def __init__(self, cfgs):
        super().__init__()
        self.cfgs = cfgs
        self.core = CamLiRAFT_L_Core(cfgs)

    def forward(self, inputs):
        """inputs: `pcs` [B,6,N], `intrinsics` [B,3] (f, cx, cy), optionally the target `flow_3d` -> {'flow_3d'}."""
        pc1, pc2 = inputs["pcs"][:, :3].float(), inputs["pcs"][:, 3:].float()
        intr = inputs["intrinsics"].float()
        persp = {"projection_mode": "perspective", "sensor_h": 540, "sensor_w": 960,
                 "f": intr[:, 0], "cx": intr[:, 1], "cy": intr[:, 2]}
        paral = None
        if self.cfgs.ids.enabled:
            qh, qw = round(540 / 32), round(960 / 32)
            paral = {"projection_mode": "parallel", "sensor_h": qh, "sensor_w": qw, "cx": (qw - 1) / 2, "cy": (qh - 1) / 2}
            pc1, pc2 = persp2paral(pc1, persp, paral), persp2paral(pc2, persp, paral)
        preds = self.core(pc1, pc2)
        if paral is not None:
            base = paral2persp(pc1, persp, paral)
            preds = [paral2persp(pc1 + p, persp, paral) - base for p in preds]
        if "flow_3d" in inputs:
            self.loss = calc_sequence_loss_3d(preds, inputs["flow_3d"][:, :3].float(), self.cfgs.loss)
        return {"flow_3d": preds[-1]}
