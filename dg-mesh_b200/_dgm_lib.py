"""ctypes binding of libdgmesh_b200.so (the C-ABI in include/dgmesh_b200.h).

PyTorch only supplies device memory (`tensor.data_ptr()`) and the CUDA stream; all
compute is in the shared library.  There is NO fallback: if the library is missing
or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgmesh_b200.so")
# experiments only: an alternative build of the same library (never a different implementation)
LIB_PATH = os.environ.get("DGMESH_B200_LIB", LIB_PATH)

c_void_p, c_int, c_float, c_size_t, c_int64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                               ctypes.c_int64)
c_int32, byref = ctypes.c_int32, ctypes.byref
P = c_void_p  # every device pointer crosses the boundary as a plain address

# name -> (restype, argtypes); must list every symbol declared in include/dgmesh_b200.h
SIGNATURES = {
    "dgm_version": (ctypes.c_char_p, []),
    "dgm_last_error": (ctypes.c_char_p, []),
    "dgr_workspace_sizes": (c_int, [c_int, c_int, c_int, c_int64, ctypes.POINTER(c_size_t),
                                    ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "dgm_notify_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "dgm_notify_host": (c_void_p, [P]),
    "dgm_notify_event": (c_void_p, [P]),
    "dgm_notify_wait": (c_int, [P]),
    "dgm_notify_destroy": (c_int, [P]),
    "dgr_forward": (c_int, [c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_float, P, P, P, P, P,
                            c_float, c_float, c_int, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, P, P,
                            c_float, c_float, c_int, P]),
    "dgr_backward": (c_int, [c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, c_float, P, P, P, P, P,
                             c_float, c_float, P, P, P, c_int64, P, P, P, P, P, P, P, P, P, P, P, P]),
    "dgr_forward_batch": (c_int, [c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_float, P, P, c_int,
                                  P, P, P, P, P, c_int, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, P, P,
                                  c_float, c_float, c_int, c_int, P]),
    "dgr_backward_batch": (c_int, [c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, c_float, P, P, c_int,
                                   P, P, P, P, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, P, P, P, P, P,
                                   P, P, P, c_int, P]),
    "dgr_mark_visible": (c_int, [c_int, P, P, P, P, P]),
    "dgr_export_state": (c_int, [c_int, c_int, c_int, c_int64, P, P, P] + [P] * 12 + [P]),
    "dgk_workspace_size": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgk_dist2": (c_int, [c_int, P, P, P, c_size_t, P]),
    "dgk_nearest": (c_int, [c_int, P, c_int, P, P, P, P]),
    "dgp_plan_create": (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_size_t)]),
    "dgp_plan_destroy": (c_int, [P]),
    "dgp_forward": (c_int, [P, c_int, ctypes.c_double, P, P, c_int, P, P, P, c_size_t, P]),
    "dgp_backward": (c_int, [P, c_int, P, P, c_int, P, P, P, P, P, c_size_t, P]),
    "dgmc_workspace_size": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgmc_count": (c_int, [c_int, P, c_float, P, c_size_t, P, P, P, P]),
    "dgmc_emit": (c_int, [c_int, P, c_float, P, c_size_t, P, c_int64, P, c_int64, P]),
    "dgmc_backward": (c_int, [c_int, c_int, P, c_float, P, c_size_t, P, P, P]),
    "dgl_gemm_ws_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "dgl_gemm_bf16": (c_int, [c_int, c_int, c_int, P, c_int, P, c_int, P, c_int, P, c_int, P, c_size_t, P]),
    "dgl_gemm_tn_bf16": (c_int, [c_int, c_int, c_int, P, c_int, P, c_int, P, c_int, c_int, P, c_size_t, P]),
    "dgl_mlp_pack_sizes": (c_int, [ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "dgl_mlp_pack": (c_int, [P, P, P, P, P]),
    "dgl_mlp_grad_pointers": (c_int, [P, P]),
    "dgl_mlp_unpack_grads": (c_int, [P, P, P, P]),
    "dgl_mlp_workspace": (c_int, [c_int, c_int, ctypes.POINTER(c_size_t)]),
    "dgl_mlp_forward": (c_int, [P, c_int, P, P, P, c_int, P, c_size_t, P]),
    "dgl_mlp_backward": (c_int, [P, c_int, P, P, P, P, c_size_t, P, P, P]),
    "dgl_laplacian_workspace": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgl_laplacian_forward": (c_int, [c_int, c_int, P, P, P, P, c_size_t, P]),
    "dgl_laplacian_backward": (c_int, [c_int, c_int, P, P, P, P, c_size_t, P]),
    "dgmr_rasterize": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "dgmr_rasterize_bwd": (c_int, [c_int, c_int, P, P, P, P, P, P]),
    "dgmr_interpolate": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
    "dgmr_interpolate_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "dgmr_antialias": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "dgmr_antialias_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, P, P]),
    "dgd_workspace_size": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgd_plan": (c_int, [c_int, P, P, P, P, c_float, c_float, c_float, c_float, c_int, c_float, P, c_size_t, P, P]),
    "dgd_split_stds": (c_int, [c_int, P, P, c_size_t, P, P]),
    "dgd_apply": (c_int, [c_int, c_int, P, P, P, P, c_size_t, P]),
    "dgx_allreduce_nvls": (c_int, [P, c_size_t, P, c_int, c_int, ctypes.c_uint32, c_float, c_int, P]),
    "dgm_profile_enable": (c_int, [c_int]),
    "dgloss_workspace_size": (c_int, [c_int, c_int, ctypes.POINTER(c_size_t)]),
    "dgloss_forward": (c_int, [c_int, c_int, P, P, c_float, c_int, P, P, c_size_t, P]),
    "dgloss_backward": (c_int, [c_int, c_int, P, P, c_float, c_int, P, P, P, c_size_t, P]),
    "dgm_profile_read": (c_int, [ctypes.POINTER(c_float), c_int]),
    "dgm_timeline_read": (c_int, [ctypes.POINTER(c_float), ctypes.POINTER(c_float), ctypes.POINTER(c_int), c_int]),
}

KERNEL_NAMES = ["preprocess", "tile_scan", "scatter", "sort_pack", "render_fwd", "render_bwd", "preprocess_bwd",
                "count"]


def profile_read():
    """{kernel name: ms of its most recent launch} (synchronises the profiling events)."""
    buf = (c_float * 16)()
    check(lib().dgm_profile_read(buf, 16), "dgm_profile_read")
    return {n: float(buf[i]) for i, n in enumerate(KERNEL_NAMES) if buf[i] >= 0}



class DglNet(ctypes.Structure):
    """Mirror of `struct DglNet` (include/dgmesh_b200.h)."""
    _fields_ = [("has_timenet", c_int), ("in_t", c_int), ("n_out", c_int), ("sigmoid_out", c_int),
                ("W", c_void_p * 8), ("WT", c_void_p * 8), ("b", c_void_p * 8),
                ("Wh", c_void_p), ("WhT", c_void_p), ("bh", c_void_p),
                ("Wt0", c_void_p), ("bt0", c_void_p), ("Wt1", c_void_p), ("Wt1T", c_void_p), ("bt1", c_void_p),
                ("precise", c_int), ("Wlo", c_void_p * 8), ("Whlo", c_void_p), ("Wt0lo", c_void_p),
                ("Wt1lo", c_void_p)]


class DglRaw(ctypes.Structure):
    """Mirror of `struct DglRaw` (reference-shaped fp32 parameters)."""
    _fields_ = [("has_timenet", c_int), ("in_t", c_int), ("sigmoid_out", c_int), ("n_heads", c_int),
                ("head_rows", c_int * 4), ("W", c_void_p * 8), ("b", c_void_p * 8), ("Wh", c_void_p * 4),
                ("bh", c_void_p * 4), ("Wt0", c_void_p), ("bt0", c_void_p), ("Wt1", c_void_p), ("bt1", c_void_p)]


class DglRawGrads(ctypes.Structure):
    _fields_ = [("W", c_void_p * 8), ("b", c_void_p * 8), ("Wh", c_void_p * 4), ("bh", c_void_p * 4),
                ("Wt0", c_void_p), ("bt0", c_void_p), ("Wt1", c_void_p), ("bt1", c_void_p)]


class DgdField(ctypes.Structure):
    """Mirror of `struct DgdField` (densify_and_prune gather)."""
    _fields_ = [("src", c_void_p), ("m1_src", c_void_p), ("m2_src", c_void_p), ("dst", c_void_p),
                ("m1_dst", c_void_p), ("m2_dst", c_void_p), ("width", c_int), ("role", c_int)]


class DglGrads(ctypes.Structure):
    """Mirror of `struct DglGrads`."""
    _fields_ = [("dW", c_void_p * 8), ("db", c_void_p * 8), ("dWh", c_void_p), ("dbh", c_void_p),
                ("dWt0", c_void_p), ("dbt0", c_void_p), ("dWt1", c_void_p), ("dbt1", c_void_p)]


_lib = None


class DgmError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgmError(
                f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
                "dgmesh_b200 has no CPU / PyTorch fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dgm_last_error().decode()
        raise DgmError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device address of a tensor, or None (NULL) for an absent / zero-element tensor
    (the reference signals "absent" with zero-element tensors, dgr/.../__init__.py:197-207)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device.  `torch.cuda.current_stream()` builds a
    Stream object (~28 us); the raw query is ~1 us -- this is called twice per rendered frame."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
