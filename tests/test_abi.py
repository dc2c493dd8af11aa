"""CPU: the C-ABI library loads and exports every symbol include/camli_b200.h declares."""
import pytest
import torch

from camliflow_b200 import native


def test_library_exports_every_declared_symbol():
    handle = native.lib()
    names = native.declared_symbols()
    assert len(names) >= 7
    for name in names:
        assert hasattr(handle, name), "missing export: " + name


def test_abi_version_and_strerror():
    handle = native.lib()
    assert handle.camli_abi_version() == native.ABI_VERSION
    assert handle.camli_strerror(0) == b"ok"
    assert b"invalid" in handle.camli_strerror(-1)


def test_argument_errors_do_not_launch():
    handle = native.lib()
    assert handle.camli_k_nearest_neighbor(1, 4, 4, 65, 3, None, None, None, None) == -2   # k > MAX_K
    assert handle.camli_k_nearest_neighbor(1, 4, 4, 0, 3, None, None, None, None) == -1
    assert handle.camli_k_nearest_neighbor(1, 4, 4, 3, 4, None, None, None, None) == -2   # D must be 2 or 3
    assert handle.camli_furthest_point_sampling(None, None, 1, 0, 1, None, None) == -1
    assert handle.camli_correlation_forward(None, None, None, 1, 0, 4, 4, 4, None) == -1


def test_host_wrappers_refuse_cpu_tensors_loudly():
    from camliflow_b200 import csrc
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        csrc.furthest_point_sampling(torch.rand(1, 10, 3), 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        csrc.k_nearest_neighbor(torch.rand(1, 10, 3), torch.rand(1, 5, 3), 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        csrc.correlation2d(torch.rand(1, 32, 4, 4), torch.rand(1, 32, 4, 4), 4)
    with pytest.raises(NotImplementedError):
        csrc.k_nearest_neighbor(torch.rand(1, 10, 3), torch.rand(1, 5, 3), 2, cpp_impl=False)
    with pytest.raises(AssertionError):
        csrc.furthest_point_sampling(torch.rand(1, 4, 3), 4)     # wrapper.py:98: N > S


def test_new_entry_points_validate_arguments_without_launching():
    """The tensor-core / small convolution and backward entry points return CAMLI_E* for bad arguments and CAMLI_OK
    for an empty batch, before any CUDA call (no GPU needed)."""
    import ctypes
    h = native.lib()
    f32, i64 = ctypes.c_float, ctypes.c_int64
    # camli_conv_gemm(x,B,H,W,Cin,ldx, w_hi,w_lo,Cout,kh,kw, bias,residual,ldr, act,slope, out,ldo, tile_n, stream)
    cg = lambda B, H, W, Cin, ldx, Cout, kh, kw, act, ldo, tile=0: h.camli_conv_gemm(       # noqa: E731
        None, B, H, W, Cin, i64(ldx), None, None, Cout, kh, kw, None, None, i64(0), act, f32(0.1), None, i64(ldo), tile, None)
    assert cg(0, 4, 4, 32, 32, 16, 3, 3, 1, 16) == 0            # empty batch
    assert cg(1, 4, 4, 32, 32, 16, 3, 3, 9, 16) == -1           # unknown activation
    assert cg(1, 4, 4, 32, 16, 16, 3, 3, 1, 16) == -1           # pixel stride smaller than Cin
    assert cg(1, 4, 4, 30, 32, 16, 3, 3, 1, 16) == -2           # Cin % 4 != 0: not TMA-tileable
    assert cg(1, 4, 4, 32, 32, 16, 2, 3, 1, 16) == -2           # even window
    assert cg(1, 4, 4, 32, 32, 16, 3, 3, 1, 16) == -1           # null pointers with a non-empty batch
    # GRU epilogues need their side inputs
    assert h.camli_conv_gemm_fused(None, 1, 4, 4, 32, i64(32), None, None, 64, 1, 1, None, None, i64(0), 5, f32(0), None, i64(64),
                                   None, i64(0), None, i64(0), 32, None, i64(0), 0, None) == -1
    sm = lambda Cin, Cout: h.camli_conv_small_n(None, 1, 4, 4, Cin, i64(Cin), None, Cout, 3, 3, None, 0, f32(0), None, i64(Cout), None)  # noqa: E731
    assert sm(32, 5) == -2 and sm(30, 2) == -2 and sm(32, 2) == -1
    assert h.camli_conv_small_cin(None, 1, 4, 4, 5, i64(5), None, 32, 3, 3, None, 0, f32(0), None, i64(32), None) == -2
    assert h.camli_corr2d_lookup_backward(None, None, None, 4, None, None, 1, 4, 4, 3, None) == -2      # radius != 4
    assert h.camli_pointconv_dw_gather_max_backward(1, 8, 4, 3, 4, 16, None, None, None, None, None, None, None) == -1   # K < k
    assert h.camli_split_tf32(None, None, None, i64(0), None) == 0
    # camli_convex_upsample(B,H,W,factor, flow, mask_rows, scale, up, stream)
    assert h.camli_convex_upsample(0, 4, 4, 8, None, None, f32(0.25), None, None) == 0            # empty batch
    assert h.camli_convex_upsample(1, 4, 4, 3, None, None, f32(0.25), None, None) == -2           # factor is 4 or 8
    assert h.camli_convex_upsample(1, 4, 4, 8, None, None, f32(0.25), None, None) == -1           # null pointers
    assert h.camli_convex_upsample_backward(1, 4, 4, 4, None, None, f32(1), None, None, None, None, None) == -1
    # camli_pointconv_group_backward(B,N,S,K,k,C, rows,ld, centre,sb,sp,sd, idx, W1,b1,W2,b2, slope, g_out, g_rows, g_centre, g_params, stream)
    pgb = lambda S, K, k, C, ld: h.camli_pointconv_group_backward(                                                  # noqa: E731
        1, 64, S, K, k, C, None, i64(ld), None, i64(0), i64(1), i64(S), None, None, None, None, None, f32(0.1), None, None, None, None, None)
    assert pgb(0, 16, 16, 35, 35) == 0 and pgb(8, 8, 16, 35, 35) == -1 and pgb(8, 16, 16, 300, 300) == -2 and pgb(8, 16, 16, 35, 35) == -1


def test_training_entry_points_validate_arguments_without_launching():
    """camli_transpose_split / camli_conv_wgrad / the single-pass flag of camli_conv_gemm: CAMLI_E* for bad arguments, CAMLI_OK
    for empty work, before any CUDA call (no GPU needed)."""
    import ctypes
    h = native.lib()
    f32, i64 = ctypes.c_float, ctypes.c_int64
    # camli_transpose_split(rows, ld, P, C, y_rows, ldy, act, slope, W, n_shift, shift_step, xstride, hi_t, lo_t, g_rows, colsum, stream)
    ts = lambda ld, P, C, act, W, n_shift, step=1, xs=1: h.camli_transpose_split(           # noqa: E731
        None, i64(ld), i64(P), C, None, i64(0), act, f32(0.1), W, n_shift, step, xs, None, None, None, None, None)
    assert ts(32, 64, 32, 0, 8, 1, 1, 3) == -1         # xstride is 1 or 2
    assert ts(32, 0, 32, 0, 8, 1) == 0                 # no rows
    assert ts(16, 64, 32, 0, 8, 1) == -1               # pitch smaller than C
    assert ts(32, 64, 32, 0, 8, 2) == -1               # even number of shifted copies
    assert ts(32, 64, 32, 5, 8, 1) == -2               # GRU epilogue codes have no derivative here
    assert ts(32, 64, 32, 0, 8, 3) == -1               # null pointers with rows to move (and P % W == 0)
    assert ts(32, 60, 32, 0, 8, 3) == -1               # shifted copies need whole image rows
    # camli_conv_wgrad(g_hi, g_lo, x_hi, x_lo, B, H, W, Cout, Cin, kh, kw, dilation, stride, Hin, passes, dw, stream)
    wg = lambda B, H, W, Cout, Cin, kh, kw, dil, passes, stride=1, Hin=None: h.camli_conv_wgrad(   # noqa: E731
        None, None, None, None, B, H, W, Cout, Cin, kh, kw, dil, stride, H if Hin is None else Hin, passes, None, None)
    assert wg(1, 4, 8, 16, 32, 3, 3, 1, 3, stride=2, Hin=4) == -1      # 4 input rows give 2 output rows at stride 2, not 4
    assert wg(0, 4, 8, 16, 32, 3, 3, 1, 3, stride=2, Hin=8) == 0
    assert wg(0, 4, 8, 16, 32, 3, 3, 1, 3) == 0        # empty batch
    assert wg(1, 4, 8, 16, 32, 3, 3, 1, 2) == -1       # passes is 1 or 3
    assert wg(1, 4, 6, 16, 32, 3, 3, 1, 3) == -2       # W % 4 != 0: rows not 16-byte granular
    assert wg(1, 4, 8, 16, 30, 3, 3, 1, 3) == -2       # Cin % 4 != 0
    assert wg(1, 4, 8, 16, 32, 2, 3, 1, 3) == -2       # even window
    assert wg(1, 4, 8, 16, 32, 3, 3, 1, 1) == -1       # null pointers with work to do
    # the single-pass flag travels in tile_n: 0x100 | width
    cg = lambda tile: h.camli_conv_gemm(None, 0, 4, 4, 32, i64(32), None, None, 16, 3, 3, None, None, i64(0), 1, f32(0.1), None,   # noqa: E731
                                        i64(16), tile, None)
    assert cg(0x100) == 0 and cg(0x100 | 64) == 0
    bad = h.camli_conv_gemm(None, 1, 4, 4, 32, i64(32), None, None, 16, 3, 3, None, None, i64(0), 1, f32(0.1), None, i64(16), 0x100 | 48, None)
    assert bad == -1                                   # 48 is not a tile width


def test_training_doorway_pads_ragged_channel_counts():
    """tc._pad_dense: input / output channel counts are padded to multiples of 4 with zeros (TMA rows are 16-byte granular),
    differentiably, and the caller slices the result back."""
    from camliflow_b200 import tc
    x = torch.randn(1, 2, 8, 147, requires_grad=True)
    w = torch.randn(126, 147, 3, 3, requires_grad=True)
    b = torch.randn(126, requires_grad=True)
    xr, w4, b4, O = tc._pad_dense(x, w, b)
    assert xr.shape == (1, 2, 8, 148) and w4.shape == (128, 148, 3, 3) and b4.shape == (128,) and O == 126
    assert float(xr[..., 147:].abs().sum()) == 0 and float(w4[126:].abs().sum()) == 0 and float(w4[:, 147:].abs().sum()) == 0
    (xr.sum() + w4.sum() + b4.sum()).backward()
    assert x.grad.shape == x.shape and w.grad.shape == w.shape and b.grad.shape == b.shape
    xr2, w42, b42, _ = tc._pad_dense(torch.randn(1, 1, 4, 64), torch.randn(32, 64, 1, 1), None)
    assert xr2.shape[-1] == 64 and w42.shape == (32, 64, 1, 1) and b42 is None
