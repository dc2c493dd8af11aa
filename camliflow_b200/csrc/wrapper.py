"""Host side of the three native extensions, mirroring models/csrc/wrapper.py.

Same names, argument meaning and error behaviour as the reference wrapper, with
two deliberate differences: (1) there is no pure-PyTorch fallback -- a CPU tensor
or a missing library raises RuntimeError instead of silently degrading
(wrapper.py:9-15,100,124 of the reference); (2) channel-first point tensors are
searched in place through strides rather than transposed and copied
(wrapper.py:119-122).
"""
import torch

from .. import native
from ..native import i32, i64, ptr, stream


# --- pybind-level callables (names of correlation.cpp:38-41, furthest_point_sampling.cpp:19-21,
# --- k_nearest_neighbor.cpp:27-29) -----------------------------------------------------------

def _furthest_point_sampling_cuda(points_xyz: torch.Tensor, n_samples: int) -> torch.Tensor:
    """furthest_point_sampling_cuda (furthest_point_sampling.cpp:5-17): [B,N,3] f32 -> [B,S] i64."""
    native.require_cuda_f32(points_xyz, "points_xyz")
    batch_size, n_points = points_xyz.shape[0], points_xyz.shape[1]
    out = torch.empty((batch_size, n_samples), dtype=torch.int64, device=points_xyz.device)
    scratch = None
    if n_points > 8192:   # only the streaming kernel needs the [B,N] running-distance buffer
        scratch = torch.empty((batch_size, n_points), dtype=torch.float32, device=points_xyz.device)
    with torch.cuda.device(points_xyz.device):
        native.call("camli_furthest_point_sampling",
                    ptr(points_xyz), ptr(scratch), i32(batch_size), i32(n_points), i32(n_samples), ptr(out), stream(),
                    algo_bytes=batch_size * (n_points * 12 + n_samples * 8), flops=batch_size * n_samples * n_points * 8)
    return out


def _k_nearest_neighbor_cuda(input_xyz: torch.Tensor, query_xyz: torch.Tensor, k: int) -> torch.Tensor:
    """k_nearest_neighbor_cuda (k_nearest_neighbor.cpp:5-24): input [B,m,D], query [B,n,D] -> [B,n,k] i64."""
    native.require_cuda_f32(input_xyz, "input_xyz")
    native.require_cuda_f32(query_xyz, "query_xyz")
    batch_size, n_queries, n_dim = query_xyz.shape
    n_inputs = input_xyz.shape[1]
    out = torch.empty((batch_size, n_queries, k), dtype=torch.int64, device=query_xyz.device)
    with torch.cuda.device(query_xyz.device):
        native.call("camli_k_nearest_neighbor",
                    i32(batch_size), i32(n_queries), i32(n_inputs), i32(k), i32(n_dim),
                    ptr(query_xyz), ptr(input_xyz), ptr(out), stream(),
                    algo_bytes=batch_size * ((n_queries + n_inputs) * n_dim * 4 + n_queries * k * 8),
                    flops=batch_size * n_queries * n_inputs * (3 * n_dim - 1))
    return out


def _k_nearest_neighbor_strided(input_view: torch.Tensor, query_view: torch.Tensor, k: int) -> torch.Tensor:
    """Same search on [B,points,D] *views* of any stride (no copy)."""
    native.require_cuda_f32(input_view, "input_xyz", contiguous=False)
    native.require_cuda_f32(query_view, "query_xyz", contiguous=False)
    batch_size, n_queries, n_dim = query_view.shape
    n_inputs = input_view.shape[1]
    out = torch.empty((batch_size, n_queries, k), dtype=torch.int64, device=query_view.device)
    qs, is_ = query_view.stride(), input_view.stride()
    with torch.cuda.device(query_view.device):
        native.call("camli_k_nearest_neighbor_strided",
                    i32(batch_size), i32(n_queries), i32(n_inputs), i32(k), i32(n_dim),
                    ptr(query_view), i64(qs[0]), i64(qs[1]), i64(qs[2]),
                    ptr(input_view), i64(is_[0]), i64(is_[1]), i64(is_[2]), ptr(out), stream(),
                    algo_bytes=batch_size * ((n_queries + n_inputs) * n_dim * 4 + n_queries * k * 8),
                    flops=batch_size * n_queries * n_inputs * (3 * n_dim - 1))
    return out


def _correlation_forward_cuda(input1: torch.Tensor, input2: torch.Tensor, max_displacement: int) -> torch.Tensor:
    """correlation_forward_cuda (correlation.cpp:11-22): NHWC f32 inputs -> [B,(2d+1)^2,H,W]."""
    native.require_cuda_f32(input1, "input1")
    native.require_cuda_f32(input2, "input2")
    batch_size, height, width, in_channels = input1.shape
    n_disp = (2 * max_displacement + 1) ** 2
    out = torch.empty((batch_size, n_disp, height, width), dtype=torch.float32, device=input1.device)
    with torch.cuda.device(input1.device):
        native.call("camli_correlation_forward",
                    ptr(out), ptr(input1), ptr(input2), i32(batch_size), i32(in_channels), i32(height), i32(width),
                    i32(max_displacement), stream(),
                    algo_bytes=batch_size * height * width * (2 * in_channels * 4 + n_disp * 4),
                    flops=2 * n_disp * in_channels * batch_size * height * width)
    return out


def _correlation_backward_cuda(grad_output, input1, input2, max_displacement: int):
    """correlation_backward_cuda (correlation.cpp:24-36): returns (grad_input1, grad_input2), both NCHW."""
    native.require_cuda_f32(input1, "input1")
    native.require_cuda_f32(input2, "input2")
    grad_output = grad_output.contiguous()
    native.require_cuda_f32(grad_output, "grad_output")
    batch_size, height, width, in_channels = input1.shape
    grad1 = torch.empty((batch_size, in_channels, height, width), dtype=torch.float32, device=input1.device)
    grad2 = torch.empty_like(grad1)
    with torch.cuda.device(input1.device):
        n_disp = (2 * max_displacement + 1) ** 2
        native.call("camli_correlation_backward",
                    ptr(grad_output), ptr(grad1), ptr(grad2), ptr(input1), ptr(input2), i32(batch_size),
                    i32(in_channels), i32(height), i32(width), i32(max_displacement), stream(),
                    algo_bytes=batch_size * height * width * (n_disp * 4 + 4 * in_channels * 4),
                    flops=4 * n_disp * in_channels * batch_size * height * width)
    return grad1, grad2


# --- public surface (models/csrc/__init__.py:1) -------------------------------------------------

class CorrelationFunction(torch.autograd.Function):
    """Same contract as the reference's CorrelationFunction (wrapper.py:18-37): NHWC inputs,
    gradients returned in NHWC."""

    @staticmethod
    def forward(ctx, input1, input2, max_displacement):
        ctx.save_for_backward(input1, input2)
        ctx.max_displacement = max_displacement
        return _correlation_forward_cuda(input1, input2, max_displacement)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad1, grad2 = _correlation_backward_cuda(grad_output, input1, input2, ctx.max_displacement)
        return grad1.permute(0, 2, 3, 1).contiguous(), grad2.permute(0, 2, 3, 1).contiguous(), None


def _no_fallback(name):
    raise NotImplementedError(
        "%s(cpp_impl=False): camliflow_b200 ships no eager-PyTorch fallback; the CUDA path is the only path" % name)


def correlation2d(input1: torch.Tensor, input2: torch.Tensor, max_displacement: int, cpp_impl=True):
    """PWC cost volume of two NCHW feature maps -> [B,(2d+1)^2,H,W] (wrapper.py:40-57)."""
    if not cpp_impl:
        _no_fallback("correlation2d")
    input1 = input1.permute(0, 2, 3, 1).contiguous().float()
    input2 = input2.permute(0, 2, 3, 1).contiguous().float()
    return CorrelationFunction.apply(input1, input2, max_displacement)


def squared_distance(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """Pairwise squared distances [B,n1,n2] of channel-last point sets (wrapper.py:60-72).
    Pure tensor algebra in the reference too (no native kernel behind it)."""
    assert xyz1.shape[-1] == xyz2.shape[-1] and xyz1.shape[-1] <= 3
    cross = torch.matmul(xyz1, xyz2.transpose(1, 2))
    sq1 = (xyz1 * xyz1).sum(-1, keepdim=True)
    sq2 = (xyz2 * xyz2).sum(-1).unsqueeze(1)
    return -2 * cross + sq1 + sq2


def furthest_point_sampling(xyz: torch.Tensor, n_samples: int, cpp_impl=True):
    """Indices [B,S] (i64) of furthest-point samples of xyz [B,N,3] (wrapper.py:75-103)."""
    assert xyz.shape[2] == 3 and xyz.shape[1] > n_samples
    if not cpp_impl:
        _no_fallback("furthest_point_sampling")
    return _furthest_point_sampling_cuda(xyz.contiguous(), n_samples)


def k_nearest_neighbor(input_xyz: torch.Tensor, query_xyz: torch.Tensor, k: int, cpp_impl=True):
    """Indices [B,n,k] (i64) of the k nearest inputs of every query; accepts [B,N,D] or
    [B,D,N] (chosen, like the reference, by shape[1] <= 3; wrapper.py:106-127)."""
    if not cpp_impl:
        _no_fallback("k_nearest_neighbor")
    if input_xyz.shape[1] <= 3:   # channel-first: search the transposed view in place
        assert query_xyz.shape[1] == input_xyz.shape[1]
        return _k_nearest_neighbor_strided(input_xyz.transpose(1, 2), query_xyz.transpose(1, 2), k)
    return _k_nearest_neighbor_strided(input_xyz, query_xyz, k)
