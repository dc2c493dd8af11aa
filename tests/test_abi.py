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
