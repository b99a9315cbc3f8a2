"""ctypes binding of libladcast_b200.so (the C ABI declared in include/ladcast_b200.h).

There is no CPU fallback: if the shared library cannot be loaded (or built with nvcc), importing the product path
fails loudly."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libladcast_b200.so")

PRECISION_BF16 = 0
PRECISION_F32 = 1


class LadcastB200Error(RuntimeError):
    pass


class DenoiserCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "in_channels", "out_channels", "cond_channels", "num_heads", "head_dim", "num_layers", "num_single_layers",
        "num_refiner_layers", "mlp_dim", "incl_time_elapsed", "precision")]


class DcaeCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "latent_channels", "out_channels", "head_dim", "n_stages", "precision")] + [
        ("stage_channels", ctypes.c_int * 8), ("stage_layers", ctypes.c_int * 8), ("stage_is_evit", ctypes.c_int * 8),
        ("in_channels", ctypes.c_int), ("enc_stage_channels", ctypes.c_int * 8), ("enc_stage_layers", ctypes.c_int * 8),
        ("enc_stage_is_evit", ctypes.c_int * 8)]


_lib = None

_vp, _i, _i64, _f, _d, _cp = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double,
                              ctypes.c_char_p)

_SIGNATURES = {
    "lc_version": ([], _i),
    "lc_last_error": ([], _cp),
    "lc_launch_count": ([], ctypes.c_longlong),
    "lc_prof_enable": ([_i], _i),
    "lc_prof_collect": ([ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_longlong)], _i),
    "lc_prof_num_classes": ([], _i),
    "lc_prof_class_name": ([_i], _cp),
    "lc_prof_collect_all": ([ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_longlong)], _i),
    "lc_debug_gemm_trace": ([_vp], _i),
    "lc_debug_attention_trace": ([_vp], _i),
    "lc_sched_scale_input": ([_vp, _vp, _i64, _f, _vp], _i),
    "lc_sched_heun_init": ([_vp, _vp, _vp, _i64, _d, _d, _vp], _i),
    "lc_sched_heun_churn": ([_vp, _vp, _vp, _i64, _d, _d, _vp], _i),
    "lc_latent_feedback": ([_vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _vp], _i),
    "lc_denoiser_create": ([ctypes.POINTER(DenoiserCfg), ctypes.POINTER(_vp)], _i),
    "lc_denoiser_destroy": ([_vp], None),
    "lc_denoiser_load": ([_vp, _cp, _vp, ctypes.POINTER(_i64), _i, _vp], _i),
    "lc_denoiser_finalize": ([_vp, _vp], _i),
    "lc_denoiser_set_geometry": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "lc_denoiser_prepare": ([_vp, _vp, _i, _vp, _i, _vp], _i),
    "lc_denoiser_forward": ([_vp, _vp, _vp, _i, _vp, _vp], _i),
    "lc_denoiser_debug_read": ([_vp, _cp, _vp, _i64, _vp], _i),
    "lc_sched_dpmpp2m_step": ([_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _f, _vp], _i),
    "lc_sched_heun_step": ([_vp, _vp, _vp, _vp, _vp, _i64, _i, _d, _d, _d, _d, _d, _vp], _i),
    "lc_gemm": ([_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "lc_attention": ([_i, _vp, _vp, _i, _i, _i, _vp], _i),
    "lc_layernorm_modulate": ([_i, _vp, _vp, _i, _i, _f, _i, _vp, _vp, _i64, _vp, _vp, _vp], _i),
    "lc_patchify": ([_i, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "lc_unpatchify_gemm": ([_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "lc_gemm_bf16out": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
}

_OPTIONAL = {
    "lc_dcae_create": ([ctypes.POINTER(DcaeCfg), ctypes.POINTER(_vp)], _i),
    "lc_dcae_destroy": ([_vp], None),
    "lc_dcae_load": ([_vp, _cp, _vp, ctypes.POINTER(_i64), _i, _vp], _i),
    "lc_dcae_finalize": ([_vp, _vp], _i),
    "lc_dcae_reserve": ([_vp, _i, _i, _i, _vp], _i),
    "lc_dcae_decode": ([_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp], _i),
    "lc_dcae_decode_ens": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp], _i),
    "lc_dcae_encode": ([_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp], _i),
    "lc_dcae_debug_read": ([_vp, _cp, _vp, _i64, _vp], _i),
    "lc_pixel_shuffle_shortcut": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "lc_pixel_unshuffle_shortcut": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "lc_sphere_conv3x3": ([_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "lc_metrics_accumulate": ([_vp, _vp, _vp, _i, ctypes.c_longlong, _i, _i, _vp, _vp, _vp], _i),
    "lc_metrics_accumulate_strided": ([_vp, ctypes.c_longlong, _vp, _vp, _i, ctypes.c_longlong, _i, _i, _vp, _vp, _vp], _i),
    "lc_enable_peer_access": ([_i], _i),
    "lc_ipc_alloc": ([ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), _vp], _i),
    "lc_ipc_open": ([_vp, ctypes.POINTER(ctypes.c_void_p)], _i),
    "lc_ipc_close": ([_vp], _i),
    "lc_ipc_free": ([_vp], _i),
    "lc_metrics_accumulate_ptrs": ([_vp, _vp, _vp, _i, ctypes.c_longlong, _i, _i, _vp, _vp, _vp], _i),
    "lc_metrics_acc": ([_vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _i, _vp, _vp, _vp], _i),
    "lc_metrics_pointwise": ([_vp, _vp, _i, ctypes.c_longlong, _i, _i, _vp, _vp, _vp, _vp], _i),
}


def load():
    """Returns the ctypes library handle, building it with nvcc first if the .so is not in the tree."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build

        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (args, res) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.argtypes = args
        fn.restype = res
    for name, (args, res) in _OPTIONAL.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
    _lib = lib
    return lib


def prof_collect():
    """{class name: {"launches", "ms", "flops", "bytes"}} of the launches recorded since lc_prof_enable(1)."""
    lib = load()
    n = lib.lc_prof_num_classes()
    ms, fl, by, ln = (_d * n)(), (_d * n)(), (_d * n)(), (ctypes.c_longlong * n)()
    check(lib.lc_prof_collect_all(ms, fl, by, ln), "lc_prof_collect_all")
    return {lib.lc_prof_class_name(i).decode(): {"launches": int(ln[i]), "ms": float(ms[i]), "flops": float(fl[i]),
                                                   "bytes": float(by[i])} for i in range(n) if ln[i]}


def check(rc, what=""):
    if rc != 0:
        msg = load().lc_last_error()
        raise LadcastB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LadcastB200Error("ladcast_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise LadcastB200Error("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def ptr_any(t):
    """Device pointer of a CUDA tensor that the callee addresses with explicit strides (may be non-contiguous)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LadcastB200Error("ladcast_b200 kernels need CUDA tensors; there is no CPU fallback")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
