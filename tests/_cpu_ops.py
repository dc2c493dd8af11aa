"""Runs the product's HOST logic on a CPU-only box: every kernel doorway of camliflow_b200.ops is
answered by its plain-PyTorch formula (tests/torch_ref.py) and the index ops by the C oracle.
Test infrastructure only -- the product itself has no such path."""
import contextlib

import torch

from oracle import camliraft_oracle as co
from tests import torch_ref as R


def _knn(input_xyz, query_xyz, k, cpp_impl=True):
    if input_xyz.shape[1] > 3:
        input_xyz, query_xyz = input_xyz.transpose(1, 2), query_xyz.transpose(1, 2)
    return co.knn(input_xyz.detach(), query_xyz.detach(), k, "kernel")       # (indices carry no gradient)


def _fps(xyz, n_samples, cpp_impl=True):
    return co.fps(xyz.detach().contiguous(), n_samples, "kernel")


def _folded(layers):
    out = []
    for c in layers:
        w, b = c.folded()
        out += [w, b]
    return out


@contextlib.contextmanager
def cpu_kernels():
    import camliflow_b200.camlipwc_core as pwc
    import camliflow_b200.camlipwc_l_core as pwl
    import camliflow_b200.camliraft_core as cc
    import camliflow_b200.camliraft_l_core as cl
    import camliflow_b200.pwc_core as pw2
    import camliflow_b200.ops as ops
    import camliflow_b200.point_conv as pc
    import camliflow_b200.utils as ut

    def knn_interpolate(input_xyz, input_feat, query_xyz, k=3):
        return R.knn_interpolate(input_xyz, input_feat, query_xyz, _knn(input_xyz, query_xyz, k))

    def backwarp_3d(xyz1, xyz2, flow12, k=3):
        return xyz2 + knn_interpolate(xyz1 + flow12, -flow12, xyz2, k)

    def corr3d_build(feat1, feat2, xyzs2, k=3):
        pyr = [torch.bmm(feat1.transpose(1, 2), feat2) / feat1.shape[1]]
        for i in range(1, len(xyzs2)):
            pyr.append(R.corr3d_pool(pyr[-1], _knn(xyzs2[i - 1], xyzs2[i], k)))
        return pyr

    def corr3d_lookup_rows(xyz1, xyzs2, pyramid, W1, b1, W2, b2):
        idxs = [_knn(x, xyz1, 16) for x in xyzs2]
        return R.corr3d_lookup(xyz1, xyzs2, pyramid, idxs, W1, b1, W2, b2).transpose(1, 2).contiguous()

    patches = {
        (ops, "_need_cuda"): lambda *a: None,
        (ops, "k_nearest_neighbor"): _knn,
        (ops, "knn_interpolate"): knn_interpolate,
        (ops, "backwarp_3d"): backwarp_3d,
        (ops, "bilinear_sample_rows"): lambda f, uv: ops.rows_of(R.bilinear_sample(f, uv)),
        (ops, "convex_upsample"): lambda flow, mask, s=8, scale=1.0: co.convex_upsample(flow.float(), (scale * mask.float()).contiguous(), s),
        (ops, "corr2d_build"): R.corr2d_build,
        (ops, "corr2d_lookup"): lambda pyr, c, r, channels_last=True: R.corr2d_lookup(pyr, c, r),
        (ops, "corr3d_build"): corr3d_build,
        (ops, "corr3d_lookup_rows"): corr3d_lookup_rows,
        (ops, "pointconv_dw_weights"): lambda xyz, s, idx, k, wn: R.pointconv_dw_weights(xyz, s, idx[:, :, :k], _folded(wn.convs)),
        (ops, "pointconv_dw_gather_max"): lambda f, w, idx, k: ops.rows_of(R.pointconv_dw_gather_max(ops.cf_of(f), w, idx[:, :, :k])),
        (ops, "pointconv_group"): lambda rows, sx, idx, k, wn, slope: R.pointconv_group(
            ops.cf_of(rows)[:, :3], ops.cf_of(rows)[:, 3:], sx, idx[:, :, :k], *_folded(wn.convs), slope),
        (ops, "clfm_interp"): lambda uv, nn, f, sn, H, W: R.clfm_interp(uv, nn, ops.cf_of(f), *_folded(sn), H, W),
        (ut, "k_nearest_neighbor"): _knn, (ut, "furthest_point_sampling"): _fps,
        (pc, "k_nearest_neighbor"): _knn, (cc, "k_nearest_neighbor"): _knn, (cl, "k_nearest_neighbor"): _knn,
        (pwc, "k_nearest_neighbor"): _knn, (pwl, "k_nearest_neighbor"): _knn,
        (pwc, "correlation2d"): R.correlation2d, (pw2, "correlation2d"): R.correlation2d,
    }
    saved = {key: getattr(*key) for key in patches}
    try:
        for (mod, name), fn in patches.items():
            setattr(mod, name, fn)
        yield
    finally:
        for (mod, name), fn in saved.items():
            setattr(mod, name, fn)
