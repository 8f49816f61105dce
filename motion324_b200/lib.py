"""ctypes binding of libm324.so (C ABI declared in include/m324.h).

There is NO fallback: if the library is missing, or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("M324_LIB", os.path.join(_HERE, "libm324.so"))   # M324_LIB: tuning experiments only


class M324Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("passes", C.c_int32),
        ("a_lo_off", C.c_int32), ("w_lo_off", C.c_int32), ("bf16", C.c_int32),
        ("bias", C.c_void_p), ("gamma", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("resid_mod", C.c_int32), ("resid_div", C.c_int64),
        ("out32", C.c_void_p), ("ldo32", C.c_int64), ("out16", C.c_void_p), ("ldo16", C.c_int64),
        ("out16_lo_off", C.c_int32), ("act", C.c_int32),
        ("qn_w", C.c_void_p), ("kn_w", C.c_void_p), ("qk_eps", C.c_float), ("qk_cols", C.c_int32),
        ("force_bn128", C.c_int32),
        ("tn", C.c_int32), ("ksplit", C.c_int32), ("accumulate", C.c_int32),
        ("aux16", C.c_void_p), ("ldaux", C.c_int64), ("aux_mode", C.c_int32), ("out16_bf16", C.c_int32), ("out_scale", C.c_float),
        ("qk_rstd", C.c_void_p), ("ld_rstd", C.c_int64),
        ("head_w", C.c_void_p), ("head_part", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_ld", C.c_int64), ("q_rows", C.c_int64),
        ("k", C.c_void_p), ("k_ld", C.c_int64),
        ("v", C.c_void_p), ("v_ld", C.c_int64), ("kv_rows", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
        ("q_batch_rows", C.c_int64), ("kv_batch_rows", C.c_int64), ("q_batch_div", C.c_int32),
        ("out", C.c_void_p), ("o_ld", C.c_int64), ("scale", C.c_float),
        ("lse", C.c_void_p), ("lse_ld", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("partial_parts", C.c_int32), ("partial_index", C.c_int32),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_ld", C.c_int64), ("q_rows", C.c_int64),
        ("k", C.c_void_p), ("k_ld", C.c_int64),
        ("v", C.c_void_p), ("v_ld", C.c_int64), ("kv_rows", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
        ("q_batch_rows", C.c_int64), ("kv_batch_rows", C.c_int64), ("q_batch_div", C.c_int32),
        ("dO", C.c_void_p), ("do_ld", C.c_int64), ("lse", C.c_void_p), ("lse_ld", C.c_int64), ("D", C.c_void_p), ("d_ld", C.c_int64),
        ("dQ", C.c_void_p), ("dq_ld", C.c_int64), ("dK", C.c_void_p), ("dk_ld", C.c_int64), ("dV", C.c_void_p), ("dv_ld", C.c_int64),
        ("scale", C.c_float),
    ]


_P, _I32, _I64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> argtypes (restype is int unless noted); must match include/m324.h
SIGNATURES = {
    "m324_version": [],
    "m324_last_error": [],
    "m324_check_device": [],
    "m324_set_tuning": [_I32, _I32],
    "m324_launch_count": [],
    "m324_gemm": [C.POINTER(GemmArgs), _P],
    "m324_attention": [C.POINTER(AttnArgs), _P],
    "m324_attention_workspace_bytes": [],
    "m324_attention_plan": [C.POINTER(AttnArgs), _I32, C.POINTER(C.c_int32)],
    "m324_attention_partial_bytes": [_I32, _I32, _I32, _I32],
    "m324_attention_merge": [C.POINTER(AttnArgs), _P],
    "m324_attention_bwd": [C.POINTER(AttnBwdArgs), _P],
    "m324_layernorm": [_P, _I64, _P, _P, _F, _I64, _I32, _I32, _I64, _I64, _P, _I64, _I32, _P, _I64, _P],
    "m324_point_embed_features": [_P, _I32, _P, _I64, _I32, _P],
    "m324_point_extra_features": [_P, _P, _I32, _P, _I64, _I32, _I32, _I32, _P],
    "m324_preprocess_frames": [_P, _I32, _I32, _I32, _I32, _P, _I64, _I32, _P],
    "m324_dino_assemble": [_P, _P, _P, _I32, _I32, _I32, _P, _P],
    "m324_assemble_tokens": [_P, _P, _P, _F, _P, _P, _P, _P, _P, _F, _I32, _I32, _I32, _I32, _I32, _P, _F, C.c_uint64, _P, _P],
    "m324_head3_mse": [_P, _I64, _P, _P, _I64, _I32, _P, _P, _P, C.POINTER(C.c_int32), _I32, _P],
    "m324_head3_from_partials": [_P, _I32, _P, _I64, _P, _P, _P, C.POINTER(C.c_int32), _P],
    "m324_mse_finalize": [_P, _I32, _D, _F, _P, _P],
    "m324_mse_loss": [_P, _P, _I64, _F, _P, _P, _P],
    "m324_cast_pad_f16": [_P, _I64, _I32, _I32, _P, _I64, _I32, _I32, _P],
    "m324_smooth_trajectories": [_P, _P, _I32, _I32, _I32, _F, _F, _I32, _I32, _P],
    "m324_filter_trajectories": [_P, _P, _I32, _I32, _I32, _I32, _P, _I32, _F, _F, _P],
    "m324_layernorm_bwd": [_P, _I64, _P, _I64, _P, _F, _I64, _I32, _I32, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _F, _P],
    "m324_qknorm_bwd": [_P, _I64, _P, _I64, _P, _I64, _P, _P, _I32, _I32, _I32, _I64, _P, _I64, _P, _P, _F, _P],
    "m324_head_bwd": [_P, _P, _P, _I64, _P, _I64, _I32, _P, _I64, _P, _P, _F, _P],
    "m324_colsum": [_P, _I64, _I64, _I32, _P, _F, _P],
    "m324_sum_groups": [_P, _I64, _I32, _I64, _I32, _I64, _I64, _I64, _I32, _F, _I32, _P, _I64, _P, _I64, _P],
    "m324_cast_transpose_f16": [_P, _I64, _I32, _I32, _P, _I64, _I32, _P],
    "m324_add_block": [_P, _I64, _I64, _I32, _F, _I32, _P, _I64, _P],
    "m324_track_points": [_P, _P, _I32, _I32, _I64, _P, _I64, _P, _P, _I32, _P, _P, _P, _P],
    "m324_sample_texture_colors": [_P, _I64, _P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P],
    "m324_scale_by_device_scalars": [_P, _I64, _P, _P, _F, _P],
    "m324_sample_albedo": [_P, _I64, _P, _I64, _P, _P, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P],
    "m324_attn_dot": [_P, _I64, _P, _I64, _I64, _I32, _P, _I64, _P],
    "m324_chamfer_nn": [_P, _I32, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "m324_chamfer_reduce": [_P, _I32, _P, _I32, _I32, _D, _P, _P],
}

_lib = None


def load():
    """Load libm324.so (once).  Raises M324Error if it has not been built: no silent fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise M324Error(f"{LIB_PATH} is missing: build it with `python -m motion324_b200.build` "
                        "(the Motion324 B200 path has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "m324_last_error" else C.c_int64 if name in ("m324_attention_workspace_bytes", "m324_attention_partial_bytes", "m324_launch_count") else C.c_int
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().m324_last_error()
        raise M324Error(f"{what} failed ({code}): {msg.decode() if msg else '?'}")
