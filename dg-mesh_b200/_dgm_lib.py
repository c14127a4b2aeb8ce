"""ctypes binding of libdgmesh_b200.so (the C-ABI in include/dgmesh_b200.h).

PyTorch only supplies device memory (`tensor.data_ptr()`) and the CUDA stream; all
compute is in the shared library.  There is NO fallback: if the library is missing
or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgmesh_b200.so")

c_void_p, c_int, c_float, c_size_t, c_int64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                               ctypes.c_int64)
P = c_void_p  # every device pointer crosses the boundary as a plain address

# name -> (restype, argtypes); must list every symbol declared in include/dgmesh_b200.h
SIGNATURES = {
    "dgm_version": (ctypes.c_char_p, []),
    "dgm_last_error": (ctypes.c_char_p, []),
    "dgr_workspace_sizes": (c_int, [c_int, c_int, c_int, c_int64, ctypes.POINTER(c_size_t),
                                    ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "dgr_forward": (c_int, [c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_float, P, P, P, P, P,
                            c_float, c_float, c_int, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, P]),
    "dgr_backward": (c_int, [c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, c_float, P, P, P, P, P,
                             c_float, c_float, P, P, P, c_int64, P, P, P, P, P, P, P, P, P, P, P, P]),
    "dgr_forward_batch": (c_int, [c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_float, P, P, P, P, P,
                                  P, P, c_int, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, c_int, P]),
    "dgr_backward_batch": (c_int, [c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, c_float, P, P, P, P, P,
                                   P, P, P, P, c_size_t, P, c_size_t, c_int64, P, c_size_t, P, P, P, P, P, P, P, P,
                                   P, c_int, P]),
    "dgr_mark_visible": (c_int, [c_int, P, P, P, P, P]),
    "dgr_export_state": (c_int, [c_int, c_int, c_int, c_int64, P, P, P] + [P] * 12 + [P]),
    "dgk_workspace_size": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgk_dist2": (c_int, [c_int, P, P, P, c_size_t, P]),
    "dgp_plan_create": (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_size_t)]),
    "dgp_plan_destroy": (c_int, [P]),
    "dgp_forward": (c_int, [P, c_int, ctypes.c_double, P, P, c_int, P, P, P, c_size_t, P]),
    "dgp_backward": (c_int, [P, c_int, P, P, c_int, P, P, P, P, P, c_size_t, P]),
    "dgmc_workspace_size": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "dgmc_count": (c_int, [c_int, P, c_float, P, c_size_t, P, P]),
    "dgmc_emit": (c_int, [c_int, P, c_float, P, c_size_t, P, P, P]),
    "dgmc_backward": (c_int, [c_int, P, c_float, P, c_size_t, P, P, P]),
    "dgl_gemm_bf16": (c_int, [c_int, c_int, c_int, P, c_int, P, c_int, P, c_int, P, c_int, c_int, P]),
    "dgm_profile_enable": (c_int, [c_int]),
    "dgm_profile_read": (c_int, [ctypes.POINTER(c_float), c_int]),
}

KERNEL_NAMES = ["preprocess", "tile_scan", "scatter", "sort_pack", "render_fwd", "render_bwd", "preprocess_bwd",
                "count"]


def profile_read():
    """{kernel name: ms of its most recent launch} (synchronises the profiling events)."""
    buf = (c_float * 16)()
    check(lib().dgm_profile_read(buf, 16), "dgm_profile_read")
    return {n: float(buf[i]) for i, n in enumerate(KERNEL_NAMES) if buf[i] >= 0}

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


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
