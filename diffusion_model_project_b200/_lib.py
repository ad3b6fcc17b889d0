"""ctypes binding of libb2d.so (include/b2d.h).  There is NO fallback: if the CUDA library is
missing or a call fails, the caller gets an exception -- the product never routes around it."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2d.so")

ABI_VERSION = 8
B2D_MAX_SEG = 6
B2D_MAX_TAPS = 27

c_void_p, c_int, c_i32, c_i64, c_u64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class ConvDesc(C.Structure):
    """struct b2d_conv_desc (include/b2d.h)."""
    _fields_ = [
        ("nseg", c_i32),
        ("in_", c_void_p * B2D_MAX_SEG),
        ("cin", c_i32 * B2D_MAX_SEG),
        ("kbase", c_i32 * B2D_MAX_SEG),
        ("N", c_i32), ("D", c_i32), ("H", c_i32), ("W", c_i32),
        ("ntaps", c_i32),
        ("tap_dz", C.c_int8 * B2D_MAX_TAPS), ("tap_dy", C.c_int8 * B2D_MAX_TAPS), ("tap_dx", C.c_int8 * B2D_MAX_TAPS),
        ("stride_h", c_i32), ("stride_w", c_i32),
        ("OH", c_i32), ("OW", c_i32),
        ("weight", c_void_p),
        ("wrows", c_i32), ("ktot", c_i32),
        ("cout", c_i32), ("nphase", c_i32),
        ("bias", c_void_p),
        ("out", c_void_p), ("out_lo", c_void_p),
        ("out_mode", c_i32),
        ("out_H", c_i32), ("out_W", c_i32),
        ("out_sy", c_i32), ("out_sx", c_i32), ("out_oy", c_i32), ("out_ox", c_i32),
        ("out_cstride", c_i32), ("out_coff", c_i32),
        ("residual", c_void_p), ("residual_lo", c_void_p),
        ("res_cstride", c_i32),
        ("stats", c_void_p),
        ("stats_cpg", c_i32),
        ("out_scale", c_void_p), ("out_mask", c_void_p),
        ("block_n", c_i32),
        ("out_f16", c_i32), ("res_f16", c_i32),
        ("tune_ksplit", c_i32),
        ("workspace", c_void_p), ("workspace_bytes", c_i64),
        ("in_stats", c_void_p), ("in_gamma", c_void_p), ("in_beta", c_void_p),
        ("in_cpg", c_i32), ("in_creal", c_i32), ("in_f16", c_i32), ("in_act", c_i32),
        ("in_eps", c_float),
        ("in_temb", c_void_p),
        ("in_temb_row", c_void_p),
        ("in_temb_row_stride", c_i32),
        ("in_temb_ncols", c_i32),
        ("in_temb_col", c_i32),
        ("tune_flags", c_i32),
        ("op_f16", c_i32),
        ("sched_kind", c_i32),
        ("sched_x", c_void_p), ("sched_noise", c_void_p), ("sched_coef", c_void_p),
        ("sched_step_idx", c_void_p), ("sched_ticket", c_void_p), ("sched_seed_dev", c_void_p),
        ("sched_seed", c_u64),
        ("sched_step_off", c_i32), ("sched_step_inc", c_i32),
        ("sched_clip", c_i32),
        ("sched_clip_lo", c_float), ("sched_clip_hi", c_float),
        ("sched_x_bf16", c_void_p), ("sched_x_bf16_lo", c_void_p),
        ("sched_bf16_stride", c_i32),
        ("reserved", c_i32 * 2),
    ]


TUNE_NO_HALO, TUNE_NO_SPLITK, TUNE_CONTIG, TUNE_STRIDED, TUNE_STREAMK, TUNE_NO_STREAMK, TUNE_PAIR, TUNE_NO_PAIR = 1, 2, 4, 8, 16, 32, 64, 128


_SIGNATURES = {
    "b2d_version": (c_int, []),
    "b2d_last_error": (C.c_char_p, []),
    "b2d_scheduler_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_float, c_float, c_void_p, c_int, c_int, c_u64, c_void_p, c_void_p, c_int, c_void_p]),
    "b2d_q_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p]),
    "b2d_conv_plan_create": (c_int, [C.POINTER(ConvDesc), C.POINTER(c_void_p)]),
    "b2d_conv_plan_destroy": (c_int, [c_void_p]),
    "b2d_conv_run": (c_int, [c_void_p, c_void_p]),
    "b2d_conv_plan_info": (c_int, [c_void_p, C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32)]),
    "b2d_conv_plan_info2": (c_int, [c_void_p, C.POINTER(c_i32)]),
    "b2d_gn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i64, c_i32, c_void_p, c_i32, c_void_p, c_void_p,
                             c_float, c_i32, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p, c_i32, c_i32, c_void_p]),
    "b2d_maxpool2x2_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p, c_i32, c_void_p]),
    "b2d_upsample2x_nearest": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_planar_to_cl": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_i32, c_void_p, c_i32, c_void_p]),
    "b2d_cl_to_planar": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_edt2d": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_bilinear_resize": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_zero": (c_int, [c_void_p, c_i64, c_void_p]),
    "b2d_gn_gn_apply": (c_int, [c_void_p, c_i32, c_void_p, c_void_p, c_i32, c_i64, c_i32, c_void_p, c_void_p, c_void_p, c_float, c_i32,
                        c_void_p, c_void_p, c_float, c_i32, c_i32, c_void_p]),
    "b2d_maxpool2x2_gn": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_float, c_i32, c_i32, c_void_p]),
    "b2d_zstack_cl": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_i32, c_void_p]),
    "b2d_zfold_combine": (c_int, [c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32,
                          c_void_p]),
    "b2d_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                              c_i32, C.c_double, c_void_p]),
    "b2d_nmse_loss": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i64, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2d_gn_silu_bwd": (c_int, [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32,
                                c_void_p, c_void_p, c_void_p, c_float, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2d_conv_wgrad": (c_int, [c_i32, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                               c_void_p, c_i32, c_i32, c_void_p]),
    "b2d_channel_sum": (c_int, [c_void_p, c_void_p, c_i32, c_i64, c_i32, c_i32, c_void_p, c_void_p]),
    "b2d_add16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i64, c_void_p]),
    "b2d_maxpool2x2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "b2d_pack_weight": (c_int, [c_void_p, c_i32, c_i32, c_i64, c_i64, c_i32, c_i64, c_i32, c_i64, c_void_p, c_void_p, c_i64, c_i32, c_i32, c_void_p]),
    "b2d_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32,
                                  c_i32, c_i32, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES.keys())

_lib: Optional[C.CDLL] = None
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


class B2DError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libb2d.so; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2DError(
                f"{LIB_PATH} is missing: build it with `python -m diffusion_model_project_b200.build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.b2d_version() != ABI_VERSION:
            raise B2DError(f"libb2d.so ABI version {l.b2d_version()} != {ABI_VERSION}; rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b2d_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == -1 else B2DError
        raise exc(f"libb2d {what} failed ({rc}): {msg}")


def call(name: str, *args, launches: int = 1) -> None:
    global launch_count
    check(getattr(lib(), name)(*args), name)
    launch_count += launches


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
