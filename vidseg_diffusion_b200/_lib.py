"""ctypes binding of libvidseg_b200.so (include/vidseg_b200.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing, or a CUDA device is
not available when a kernel is requested, the call raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvidseg_b200.so")

c_int = ctypes.c_int
c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p
c_float = ctypes.c_float
c_longlong = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/vidseg_b200.h one to one
SIGNATURES = {
    "vidseg_last_error": (ctypes.c_char_p, []),
    "vidseg_abi_version": (c_int, []),
    "vidseg_device_arch": (c_int, []),
    "vidseg_launch_count": (c_longlong, []),
    "vidseg_profile_enable": (c_int, [c_int]),
    "vidseg_profile_read": (c_int, [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_longlong), ctypes.POINTER(ctypes.c_double)]),
    "vidseg_aggregate_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vidseg_aggregate_normalize_rows": (c_int, [c_void_p, c_int, c_longlong, c_longlong, c_int, c_void_p, c_void_p]),
    "vidseg_kmeans_exchange_words": (c_size_t, [c_void_p, c_size_t]),
    "vidseg_kmeans_exchange_mode": (c_int, [c_void_p, c_size_t, c_int, c_int]),
    "vidseg_kmeans_partial_words": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vidseg_kmeans_update_words": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_int, c_void_p]),
    "vidseg_kmeans_flags_async": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "vidseg_kmeans_workspace_bytes": (c_size_t, [c_int] * 5),
    "vidseg_kmeans_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "vidseg_kmeans_seed": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidseg_kmeans_assign": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "vidseg_kmeans_partial": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vidseg_kmeans_update": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p]),
    "vidseg_kmeans_lloyd": (c_int, [c_void_p, c_size_t, c_int, c_void_p]),
    "vidseg_kmeans_active_runs": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_int), c_void_p]),
    "vidseg_kmeans_status": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_void_p]),
    "vidseg_kmeans_inertia": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p]),
    "vidseg_kmeans_same_matrix": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p]),
    "vidseg_kmeans_pick_best_host": (c_int, [c_void_p, c_void_p, c_int]),
    "vidseg_kmeans_finish": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p]),
    "vidseg_kmeans_select": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vidseg_kmeans_predict": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vidseg_kmeans_fit_predict": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidseg_kmeans_release": (c_int, [c_void_p]),
    "vidseg_majority_map": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vidseg_knn_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vidseg_knn_predict": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "vidseg_split_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_void_p]),
    "vidseg_gemm_split": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_float, c_void_p]),
    "vidseg_gemm_split_ex": (c_int, [c_void_p] * 7 + [c_longlong, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vidseg_gemm_split_seg": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                      c_void_p]),
    "vidseg_gemm_geglu_split": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_float, c_void_p]),
    "vidseg_set_operand_mode": (c_int, [c_int]),
    "vidseg_get_operand_mode": (c_int, []),
    "vidseg_sampler_step": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "vidseg_set_kmeans_mstep": (c_int, [c_int]),
    "vidseg_get_kmeans_mstep": (c_int, []),
    "vidseg_split_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_int, c_void_p]),
    "vidseg_conv_temporal_split": (c_int, [c_void_p] * 12 + [c_int] * 5 + [c_float, c_void_p]),
    "vidseg_temporal_attention": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_float, c_void_p]),
    "vidseg_layernorm_bias_split": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                            c_longlong, c_int, c_void_p]),
    "vidseg_conv2d_split": (c_int, [c_void_p] * 10 + [c_int] * 7 + [c_float, c_void_p]),
    "vidseg_conv2d_down_pad_after_split": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float, c_void_p]),
    "vidseg_softmax_rows_split": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "vidseg_layernorm_split": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "vidseg_geglu_split": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "vidseg_groupnorm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vidseg_groupnorm_split": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vidseg_upsample2x_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vidseg_attention_split": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vidseg_segmap_difference": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vidseg_lanczos_masks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vidseg_segmap_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, ctypes.c_double, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "vidseg_refine_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vidseg_refine_masks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
}

_lib = None
_lock = threading.Lock()


class VidsegError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VidsegError(
                f"{LIB_PATH} is missing: build it with `python -m vidseg_diffusion_b200.build` "
                "(there is no CPU/PyTorch fallback for the hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(code, what=""):
    if code != 0:
        msg = load().vidseg_last_error().decode(errors="replace")
        raise VidsegError(f"{what or 'libvidseg_b200'} failed with code {code}: {msg}")


KERNEL_FAMILIES = ("other", "gemm", "attention", "conv", "aggregate", "kmeans", "refine", "elementwise")


def profile_enable(on=True):
    check(load().vidseg_profile_enable(1 if on else 0), "profile_enable")


def profile_read():
    """{family: {"ms": total ms, "launches": n, "work": FLOPs or bytes}} since the last profile_enable(True)."""
    lib = load()
    out = {}
    for i, name in enumerate(KERNEL_FAMILIES):
        ms, n, w = ctypes.c_double(), c_longlong(), ctypes.c_double()
        check(lib.vidseg_profile_read(i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)), "profile_read")
        out[name] = {"ms": ms.value, "launches": n.value, "work": w.value}
    return out


def launch_count():
    return int(load().vidseg_launch_count())


def require_cuda_tensor(t, dtype, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise VidsegError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise VidsegError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise VidsegError(f"{name}: expected a contiguous tensor")
    return t


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
