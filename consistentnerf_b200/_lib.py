"""ctypes binding of libcnerf.so (the C ABI declared in include/cnerf.h).

The library is the product: if it is missing this module raises -- there is no eager
PyTorch or CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNERF_LIB") or os.path.join(_PKG, "libcnerf.so")      # CNERF_LIB: A/B runs against another build

_P = c_void_p
_I = c_int
_F = c_float

# name -> (restype, argtypes); mirrors include/cnerf.h line by line
SIGNATURES = {
    "cnerf_version": (_I, []),
    "cnerf_last_error": (_I, [c_char_p, _I]),
    "cnerf_device_info": (_I, [POINTER(c_int), POINTER(c_int)]),
    "cnerf_pack_rays": (_I, [_P, _P, _I, _F, _F, _I, _I, _I, _I, _F, _P, _P]),
    "cnerf_image_rays": (_I, [_I, _I, POINTER(c_float), POINTER(c_float), _F, _F, _I, _I, _P, _P]),
    "cnerf_gather_rays": (_I, [_I, _I, POINTER(c_float), POINTER(c_float), _P, _I, _F, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cnerf_stratified_z": (_I, [_P, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    "cnerf_ray_points": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "cnerf_posenc": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "cnerf_linear_fwd": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "cnerf_linear_bwd_data": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _I, _P]),
    "cnerf_linear_bwd_weight_workspace": (c_int64, [_I, _I, _I]),
    "cnerf_linear_bwd_weight": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "cnerf_weights_create": (_I, [POINTER(c_void_p)]),
    "cnerf_weights_destroy": (None, [_P]),
    "cnerf_weights_refresh": (_I, [_P, POINTER(c_void_p), POINTER(c_void_p), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cnerf_mlp_fwd": (_I, [_P, _P, _P, _I, _I, _P, _I, _P]),
    "cnerf_mlp_acts_bytes": (c_int64, [c_int64]),
    "cnerf_mlp_fwd_train": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _P]),
    "cnerf_mlp_grads_bytes": (c_int64, [c_int64]),
    "cnerf_mlp_bwd_workspace_bytes": (c_int64, []),
    "cnerf_mlp_bwd": (_I, [_P, _P, _P, _P, _I, POINTER(c_void_p), POINTER(c_void_p), _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "cnerf_mlp_bwd_data": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "cnerf_mlp_bwd_weights": (_I, [_P, _P, _I, POINTER(c_void_p), POINTER(c_void_p), _P, _P, _P, _P, _I, _I, _P, _P]),
    "cnerf_mlp_bwd_heads": (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _P, _P]),
    "cnerf_composite_fwd": (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cnerf_composite_bwd": (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "cnerf_sample_pdf": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "cnerf_sample_fine": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "cnerf_project_gather": (_I, [_P, _I, POINTER(c_float), POINTER(c_float), POINTER(c_float), _P, _I, _P, _I, _I,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cnerf_hard_mask_pair": (_I, [_P, _P, _P, _I, POINTER(c_float), POINTER(c_float), _P, _I, _I, _F, _I, _I, _P, _P]),
    "cnerf_masked_mse_fwd": (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P]),
    "cnerf_masked_mse_bwd": (_I, [_P, _P, _P, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P]),
    "cnerf_soft_mse_fwd": (_I, [_P, _P, c_int64, _F, _I, _F, _P, _P, _P, _P]),
    "cnerf_soft_mse_bwd": (_I, [_P, _P, c_int64, _F, _I, _P, _P, _P, _P]),
}

# include/cnerf_debug.h: self-tests, microbenchmarks, profiling hooks -- not part of the drop-in boundary
DEBUG_SIGNATURES = {
    "cnerf_umma_selftest": (_I, [_P, _P, _I, _I, _P, _P]),
    "cnerf_umma_selftest_ts": (_I, [_P, _P, _I, _I, _P, _P]),
    "cnerf_debug_mlp_fwd_terms": (_I, [_P, _P, _P, _I, _I, _P, _I, _P]),
    "cnerf_debug_profile3": (_I, [_I, POINTER(ctypes.c_ulonglong)]),
    "cnerf_debug_profile5": (_I, [_I, POINTER(ctypes.c_ulonglong)]),
    "cnerf_debug_profile_chain": (_I, [_I, POINTER(ctypes.c_ulonglong)]),
    "cnerf_debug_umma_rate": (_I, [_I, _I, _I, _I, _P, _P]),
}
# only in builds made with `python -m consistentnerf_b200.build --experiments` (csrc/experiments/)
EXPERIMENT_SIGNATURES = {
    "cnerf_umma_selftest_pair": (_I, [_P, _P, _I, _I, _P, _P]),
    "cnerf_debug_profile4": (_I, [_I, POINTER(ctypes.c_ulonglong)]),
    "cnerf_debug_pair_layout": (_I, [_P, _P, _I, _I, _P, _P]),
    "cnerf_debug_umma_rate_pair": (_I, [_I, _I, _I, _P, _P, _P]),
}

_dll = None


def load() -> ctypes.CDLL:
    """dlopen libcnerf.so and declare every prototype.  Raises if the library is absent."""
    global _dll
    if _dll is not None:
        return _dll
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is required (no fallback path exists). "
            "Build it with `python -m consistentnerf_b200.build`.")
    dll = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
        fn = getattr(dll, name)      # AttributeError here == missing export
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in EXPERIMENT_SIGNATURES.items():
        if hasattr(dll, name):
            fn = getattr(dll, name)
            fn.restype = res
            fn.argtypes = args
    _dll = dll
    return dll


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().cnerf_last_error(buf, 512)
    return buf.value.decode(errors="replace")


# kernels launched per successful call (1 unless listed): the bench's gpu_launches claim is counted here
LAUNCHES_PER_CALL = {"cnerf_weights_refresh": 3, "cnerf_mlp_bwd": 7, "cnerf_mlp_bwd_data": 2, "cnerf_mlp_bwd_weights": 3, "cnerf_mlp_bwd_heads": 2, "cnerf_masked_mse_fwd": 2, "cnerf_soft_mse_fwd": 2, "cnerf_linear_bwd_weight": 4}
launch_count = 0
# name -> list of (start, end) CUDA event pairs; filled only for the names put into the dict by a profiler
event_trace = {}


def call(name: str, *args):
    """Invoke an int-status entry point; non-zero status raises RuntimeError(last_error)."""
    global launch_count
    trace = event_trace.get(name)
    if trace is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {last_error()}")
    if trace is not None:
        e1.record()
        trace.append((e0, e1))
    launch_count += LAUNCHES_PER_CALL.get(name, 1)


def ptr(t):
    """Device pointer of a contiguous fp32 (or given dtype) CUDA tensor, None -> NULL."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("consistentnerf_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("tensor passed to the C ABI must be contiguous")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def host_floats(values):
    """Small host matrix (row-major) as a float* argument."""
    flat = [float(v) for v in values]
    return (c_float * len(flat))(*flat)
