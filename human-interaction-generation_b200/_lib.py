"""ctypes binding of csrc/libhig_b200.so (the C ABI declared in include/hig_b200.h).

There is deliberately no fallback: if the shared library is missing it is built in-tree with nvcc, and if that
fails, or a call returns an error code, a RuntimeError is raised.  Nothing here touches the CPU oracle.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhig_b200.so")

_lib = None
_lock = threading.Lock()

c_int, c_void_p, c_float_p = ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p
c_ull = ctypes.c_ulonglong

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "hig_version": [],
    "hig_last_error": [],
    "hig_launch_count": [],
    "hig_debug_trace": [c_void_p, c_int],
    "hig_debug_saturation": [c_void_p],
    "hig_set_sm_limit": [c_int],
    "hig_l2_persist": [c_void_p, c_ull, ctypes.c_float, c_void_p],
    "hig_gemm_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                      c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "hig_gemm_bf16_ex": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                         c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "hig_gemm_stream": [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_int, c_void_p, c_int, c_void_p],
    "hig_row_stats": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "hig_gemm_f32": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                     c_void_p, c_int, c_int, c_void_p],
    "hig_ln_film_silu": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                         c_int, c_void_p],
    "hig_eff_attn": [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                     c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hig_attn_apply_stylize": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                               c_int, c_void_p],
    "hig_attn_kv": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hig_attn_apply_stylize_tc": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                  c_int, c_int, c_void_p],
    "hig_attn_apply_stylize_y": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_void_p],
    "hig_timestep_embed": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p],
    "hig_time_table_silu": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p],
    "hig_tile_rows": [c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p],
    "hig_pack_motion": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "hig_ddpm_step": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ull,
                      c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "hig_recover_joints": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                           c_void_p, c_void_p],
    "hig_q_sample": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    # training path
    "hig_gemm_bf16_splitk": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "hig_gemm_bf16_t": [c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                        c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "hig_gemm_bf16_fused": [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                            c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "hig_transpose": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                      c_void_p],
    "hig_colsum": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "hig_act_fwd": [c_void_p, c_int, ctypes.c_longlong, c_int, c_void_p, c_int, c_void_p],
    "hig_act_bwd": [c_void_p, c_int, c_void_p, c_int, ctypes.c_longlong, c_int, c_void_p, c_int, c_void_p],
    "hig_ln_film_silu_bwd": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                             c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "hig_eff_attn_bwd": [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                         c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hig_eff_attn_bwd_sums": [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_void_p, c_void_p],
    "hig_mha_attention": [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hig_masked_mse": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_void_p],
    "hig_sumsq": [c_void_p, ctypes.c_longlong, c_void_p, c_void_p],
    "hig_mean_slices": [c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_int, ctypes.c_float, c_void_p],
    "hig_adam_flat": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_float, ctypes.c_float,
                      ctypes.c_float, ctypes.c_float, c_int, c_void_p, ctypes.c_float, c_void_p],
}
_RESTYPE = {"hig_last_error": ctypes.c_char_p, "hig_launch_count": c_ull}


def ensure_built():
    if not os.path.exists(LIB_PATH):
        from .build import build
        build()
    return LIB_PATH


def load():
    """Load (building first if needed) the shared library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = ensure_built()
        lib = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library drift: fail loudly
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().hig_last_error()
        raise RuntimeError(f"hig_b200: {what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count():
    return int(load().hig_launch_count())
