"""ctypes binding of libcamli_b200.so -- the only doorway from Python to the kernels.

There is deliberately NO fallback: if the library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used for device memory and streams only; the
entry points receive raw device pointers, sizes and the current CUDA stream.
"""
import ctypes
import os
import re

import torch

from .build import LIB_PATH

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "camli_b200.h")
_lib = None

ABI_VERSION = 1


def declared_symbols(header_path=_HEADER):
    """Names of every function include/camli_b200.h declares."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(camli_\w+)\s*\(", text)))


def lib():
    """Load (once) and return the shared library; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libcamli_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m camliflow_b200.build`; there is no CPU / PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        handle.camli_strerror.restype = ctypes.c_char_p
        handle.camli_strerror.argtypes = [ctypes.c_int]
        if handle.camli_abi_version() != ABI_VERSION:
            raise RuntimeError("libcamli_b200.so ABI %d != expected %d: rebuild the library"
                               % (handle.camli_abi_version(), ABI_VERSION))
        if os.environ.get("CAMLI_PDL") is not None:         # A/B switch for the benchmarks (default: on)
            handle.camli_conv_gemm_set_pdl(int(os.environ["CAMLI_PDL"]))
        if os.environ.get("CAMLI_FPS_PATH") is not None:    # 3 pruned single CTA (default) / 2 cluster + st.async / 1 / 0
            handle.camli_fps_set_cluster_path(int(os.environ["CAMLI_FPS_PATH"]))
        if os.environ.get("CAMLI_LOOKUP_KEEP_L2") is not None:
            handle.camli_corr2d_lookup_set_l2_keep(int(os.environ["CAMLI_LOOKUP_KEEP_L2"]))
        _lib = handle
    return _lib


_launches = 0


def launch_count():
    """Number of C-ABI kernel entry points called so far in this process."""
    return _launches


def check(code, what):
    global _launches
    _launches += 1
    if code != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (what, lib().camli_strerror(code).decode(), code))


_profile = None   # name -> [(start_event, end_event, algorithmic_bytes, flops)] while profiling


def call(name, *args, algo_bytes=0, flops=0, shape=None):
    """Invoke entry point `name` of the library on the current stream; raise on a non-zero
    return code.  While profiling (profile_begin/profile_end) every call is bracketed by CUDA
    events recorded on the stream it launches on."""
    fn = getattr(lib(), name)
    if _profile is None:
        check(fn(*args), name)
        return
    st = torch.cuda.current_stream()
    # inside a stream capture the events become event-record NODES of the graph (external=True): after a replay they
    # time the kernel as it runs in the graph, back to back with its neighbours -- not an eager launch on an idle GPU
    ext = torch.cuda.is_current_stream_capturing()
    s, e = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
    s.record(st)
    code = fn(*args)
    e.record(st)
    _profile.setdefault(name, []).append((s, e, algo_bytes, flops, shape))
    check(code, name)


def profile_begin():
    global _profile
    _profile = {}


def profile_end():
    """Stops profiling; returns {name: {launches, avg_us, total_us, bytes, flops}} (per launch averages)."""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, recs in (prof or {}).items():
        us = [r[0].elapsed_time(r[1]) * 1e3 for r in recs]
        out[name] = {"launches": len(recs), "avg_us": sum(us) / len(us), "total_us": sum(us),
                     "bytes": sum(r[2] for r in recs) / len(recs), "flops": sum(r[3] for r in recs) / len(recs)}
        groups = {}                                     # the same per problem size (algorithmic bytes, flops)
        for t, (_, _, by, fl, shp) in zip(us, recs):
            groups.setdefault((by, fl, shp), []).append(t)
        out[name]["by_size"] = [{"bytes": by, "flops": fl, "shape": shp, "launches": len(ts), "avg_us": sum(ts) / len(ts),
                                 "min_us": min(ts)}
                                for (by, fl, shp), ts in sorted(groups.items(), key=lambda kv: kv[0][:2], reverse=True)]
    return out


def ptr(t):
    """Device pointer of a tensor as a ctypes void* (None -> NULL)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    """The current torch CUDA stream as a ctypes void*."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda_f32(t, name, contiguous=True):
    """Argument checks with the reference's TORCH_CHECK wording (k_nearest_neighbor.cpp:6-12 etc.)."""
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)
    if contiguous and not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


i32 = ctypes.c_int
i64 = ctypes.c_int64
