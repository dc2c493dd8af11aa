"""CamLiRAFT-L, the LiDAR-only model (reference models/camliraft_l.py:7-79; BASELINE config 1): inverse-depth
scaling of the two clouds, the point branch alone (CamLiRAFT_L_Core), and the way back to metric flow."""
from .base import FlowModel
from .camliraft_l_core import CamLiRAFT_L_Core
from .ids import paral2persp, persp2paral
from .losses import calc_sequence_loss_3d


class CamLiRAFT_L(FlowModel):
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
        transfer = "src_mean" in inputs and "dst_mean" in inputs
        if transfer:          # statistics transfer of the clouds to the training distribution (camliraft_l.py:38-58)
            src_mean, dst_mean = inputs["src_mean"][..., None].float(), inputs["dst_mean"][..., None].float()
            src_std, dst_std = inputs["src_std"][..., None].float(), inputs["dst_std"][..., None].float()
            pc1 = ((pc1 - src_mean) / src_std) * dst_std + dst_mean
            pc2 = ((pc2 - src_mean) / src_std) * dst_std + dst_mean
        preds = self.core(pc1, pc2)
        if transfer:
            back = lambda p: ((p - dst_mean) / dst_std) * src_std + src_mean     # noqa: E731
            preds = [back(pc1 + p) - back(pc1) for p in preds]
            pc1 = back(pc1)
        if paral is not None:
            base = paral2persp(pc1, persp, paral)
            preds = [paral2persp(pc1 + p, persp, paral) - base for p in preds]
        if "flow_3d" in inputs:
            target_3d = inputs["flow_3d"][:, :3].float()
            self.loss = calc_sequence_loss_3d(preds, target_3d, self.cfgs.loss)
            self.update_metrics("loss3d", self.loss)
            self.update_3d_metrics(preds[-1], target_3d)
        return {"flow_3d": preds[-1]}

    @staticmethod
    def is_better(curr_summary, best_summary):
        return best_summary is None or curr_summary["epe3d"] < best_summary["epe3d"]
