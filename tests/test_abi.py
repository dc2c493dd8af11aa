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
