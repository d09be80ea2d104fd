"""ctypes binding of libxrd.so (include/xrd.h).

There is deliberately no fallback of any kind: if the shared library is not
built, or no sm_100 GPU is present, the model classes raise.  The oracle under
``oracle/`` is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxrd.so")

XRD_MAX_LEVELS = 8
MODE_BF16, MODE_FP32_CHECK, MODE_FP16 = 0, 1, 2
PART_UNET, PART_NAFNET, PART_ROUTER, PART_FUSION, PART_ALL = 1, 2, 4, 8, 15
PART_EXPERT = 16
API_VERSION = 3

_I8 = C.c_int32 * XRD_MAX_LEVELS


class XrdConfig(C.Structure):
    """Mirror of ``xrd_config`` (include/xrd.h)."""
    _fields_ = [
        ("unet_in_channels", C.c_int32),
        ("unet_model_channels", C.c_int32),
        ("unet_n_levels", C.c_int32),
        ("unet_channel_mult", _I8),
        ("unet_num_res_blocks", C.c_int32),
        ("unet_n_attn", C.c_int32),
        ("unet_attention_resolutions", _I8),
        ("unet_time_emb_dim", C.c_int32),
        ("unet_num_heads", C.c_int32),
        ("naf_img_channel", C.c_int32),
        ("naf_width", C.c_int32),
        ("naf_middle_blk_num", C.c_int32),
        ("naf_n_enc", C.c_int32),
        ("naf_enc_blk_nums", _I8),
        ("naf_n_dec", C.c_int32),
        ("naf_dec_blk_nums", _I8),
        ("router_base_c", C.c_int32),
        ("fusion_base_c", C.c_int32),
        ("noise_steps", C.c_int32),
        ("beta_start", C.c_float),
        ("beta_end", C.c_float),
        ("unet_prefix", C.c_char * 64),
        ("naf_prefix", C.c_char * 64),
        ("router_prefix", C.c_char * 64),
        ("fusion_prefix", C.c_char * 64),
        ("expert_base_c", C.c_int32),
        ("expert_prefix", C.c_char * 64),
    ]


# every symbol include/xrd.h declares: (restype, argtypes)
_P = C.c_void_p
_F = C.c_void_p   # float* passed as raw device addresses
SYMBOLS = {
    "xrd_default_config": (None, [C.POINTER(XrdConfig)]),
    "xrd_api_version": (C.c_int, []),
    "xrd_last_error": (C.c_char_p, []),
    "xrd_kernel_launch_count": (C.c_uint64, []),
    "xrd_create": (C.c_int, [C.c_int, C.POINTER(XrdConfig), C.POINTER(_P)]),
    "xrd_destroy": (None, [_P]),
    "xrd_set_param": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "xrd_finalize_weights": (C.c_int, [_P, C.c_int]),
    "xrd_profile_begin": (C.c_int, []),
    "xrd_profile_end": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "xrd_export_weights": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "xrd_import_weights": (C.c_int, [_P, _P, C.c_uint64]),
    "xrd_set_mode": (C.c_int, [_P, C.c_int]),
    "xrd_get_mode": (C.c_int, [_P]),
    "xrd_set_use_graph": (C.c_int, [_P, C.c_int]),
    "xrd_set_side_branches": (C.c_int, [_P, C.c_int]),
    "xrd_unet_eps": (C.c_int, [_P, _F, _F, _P, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_ddim_denoise": (C.c_int, [_P, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_ddim_num_evals": (C.c_int, [C.c_int, C.c_int]),
    "xrd_nafnet": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_router": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_fusion": (C.c_int, [_P, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_hybrid": (C.c_int, [_P, _F, C.c_int, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_expert": (C.c_int, [_P, _F, _F, C.c_int, C.c_int, C.c_int, _P]),
    "xrd_set_range_audit": (C.c_int, [_P, C.c_int]),
    "xrd_get_range_report": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_float),
                                       C.POINTER(C.c_uint32), C.c_int]),
    "xrd_tiles_plan": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4 + [C.c_int]),
    "xrd_tiles_extract": (C.c_int, [_F, _F] + [C.c_int] * 5 + [_P]),
    "xrd_tiles_blend": (C.c_int, [_F, _F] + [C.c_int] * 5 + [_P]),
    "xrd_resize_bicubic_u8": (C.c_int, [_P, _P, _P] + [C.c_int] * 5 + [_P]),
    "xrd_resample_table": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "xrd_u8_to_unit": (C.c_int, [_P, _F, C.c_int64, _P]),
    "xrd_unit_to_u8": (C.c_int, [_F, _P, C.c_int64, _P]),
    "xrd_op_conv2d": (C.c_int, [_P, C.c_int, _F, _F, _F, _F] + [C.c_int] * 8 + [_P]),
    "xrd_op_conv2d_stats": (C.c_int, [_P, C.c_int, _F, _F, _F, _F, _P] + [C.c_int] * 8 + [_P]),
    "xrd_op_groupnorm_act": (C.c_int, [_P, _F, _F, _F, _F] + [C.c_int] * 6 + [_P]),
    "xrd_op_attention": (C.c_int, [_P, C.c_int, _F, _F] + [C.c_int] * 5 + [_P]),
    "xrd_op_time_last": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
}

_lib = None
_lock = threading.Lock()


def loaded() -> bool:
    """True once libxrd.so has been dlopen'ed by this process (never triggers the load)."""
    return _lib is not None


class XrdError(RuntimeError):
    """Raised for any non-zero status from libxrd (RUN:96-101 maps it to a
    null result for that model, RUN:210-213 to HTTP 500, as for a PyTorch error)."""


def load() -> C.CDLL:
    """dlopen libxrd.so and type every entry point.  Fails loudly when absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise XrdError(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)           # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.xrd_api_version() != API_VERSION:
            raise XrdError("libxrd.so API version mismatch; rebuild")
        _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().xrd_last_error()
        raise XrdError(f"libxrd error {status}: {msg.decode() if msg else '?'}")


def default_config() -> XrdConfig:
    cfg = XrdConfig()
    load().xrd_default_config(C.byref(cfg))
    return cfg


def profile_begin() -> None:
    """Start the in-situ launch profile of the calling thread (include/xrd.h: xrd_profile_begin).  Eager launches only: set
    ``model.use_cuda_graph = False`` for the profiled calls."""
    check(load().xrd_profile_begin())


def profile_end() -> dict:
    """Stop it; returns {kernel: (launches, total_ms)} in order of first launch."""
    lib = load()
    need = C.c_uint64(0)
    buf = C.create_string_buffer(1 << 16)
    check(lib.xrd_profile_end(buf, len(buf), C.byref(need)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split("\t")
        out[name] = (int(n), float(ms))
    return out
