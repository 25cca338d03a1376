"""ctypes binding of libgaitk.so (include/gaitk.h).  No torch types cross this boundary: tensors are
passed as raw device pointers + sizes and the current CUDA stream handle.

The library is required: there is NO Python / eager / CPU fallback for any compute entry.  Loading
works without a GPU (symbol table only); creating a plan on a non-sm_100 device fails with
GAITK_E_ARCH.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("GAITK_LIB", _HERE / "libgaitk.so"))

MAX_STREAMS, MAX_CLASSES = 3, 4
FAMILY_WEARGAIT, FAMILY_FOG = 0, 1
DTYPE_F32, DTYPE_TF32, DTYPE_BF16X3 = 0, 1, 2
SOLVER_SLSQP, SOLVER_EXACT, SOLVER_MEAN = 0, 1, 2
DENOM_FLOATS = 32           # GAITK_DENOM_FLOATS
SOLVER_FLAG_CHECK_EXCHANGE, DIAG_FLOATS, DIAG_EXCHANGE = 0x100, 24, 16

EXPORTS = [
    "gaitk_version", "gaitk_last_error", "gaitk_plan_create", "gaitk_plan_destroy", "gaitk_param_count",
    "gaitk_param_info", "gaitk_param_total", "gaitk_shared_total", "gaitk_num_streams", "gaitk_stream_in_dim",
    "gaitk_stream_in_len", "gaitk_stream_geometry", "gaitk_workspace_bytes", "gaitk_forward", "gaitk_loss", "gaitk_backward",
    "gaitk_step_grads", "gaitk_gbuf_floats", "gaitk_loss_denominators", "gaitk_step_update", "gaitk_p2p_allreduce", "gaitk_cagrad", "gaitk_cagrad_solve_host",
    "gaitk_sgd", "gaitk_window_indices", "gaitk_stats_accumulate", "gaitk_stats_finalize",
    "gaitk_normalize_frames", "gaitk_window_gather", "gaitk_mask_eval", "gaitk_fog_prepare_pose", "gaitk_fog_prepare_sensor",
    "gaitk_umma_selftest", "gaitk_umma_selftest_bf16", "gaitk_umma_bench", "gaitk_umma_bench_multi",
    "gaitk_stage_create", "gaitk_stage_forward", "gaitk_stage_backward", "gaitk_xattn_forward", "gaitk_xattn_backward",
    "gaitk_linear_forward", "gaitk_linear_workspace_bytes", "gaitk_linear_backward", "gaitk_adam",
]


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "family", "T", "enc_out_ch", "shared_out_ch", "backbone_dim", "num_classes", "use_norm", "use_cosine",
        "synchronized", "skel_in_dim", "sensor_in_ch", "sensor_len", "sensor_out_len")] + [("reserved", C.c_int32 * 3)]


class StageDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("enc", "CIN", "H", "C", "T_in", "T", "pool_sensor", "S", "bdim")] + [("reserved", C.c_int32 * 7)]


STAGE_CONV_GELU_LN, STAGE_INSOLE, STAGE_LINEAR_LN_RELU, STAGE_CONV_POOL, STAGE_TRUNK = 0, 1, 2, 3, 4


class LossDesc(C.Structure):
    _fields_ = [("scale", C.c_float), ("margin", C.c_float * MAX_CLASSES), ("cls_weight", C.c_float * MAX_CLASSES),
                ("nan_if_degenerate", C.c_int32), ("reserved", C.c_int32 * 2)]


class GaitkError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libgaitk.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GaitkError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a).  gaitk has no CPU or eager fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u32, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_float, C.c_size_t
    pp = C.POINTER(C.c_void_p)
    L.gaitk_version.restype = i32
    L.gaitk_last_error.restype = C.c_char_p
    L.gaitk_plan_create.argtypes = [C.POINTER(ModelDesc), i32, pp]; L.gaitk_plan_create.restype = i32
    L.gaitk_plan_destroy.argtypes = [vp]; L.gaitk_plan_destroy.restype = None
    L.gaitk_param_count.argtypes = [vp]; L.gaitk_param_count.restype = i32
    L.gaitk_param_info.argtypes = [vp, i32, C.c_char_p, sz, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32)]
    L.gaitk_param_info.restype = i32
    for f in ("gaitk_param_total", "gaitk_shared_total", "gaitk_gbuf_floats"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = i64
    L.gaitk_num_streams.argtypes = [vp]; L.gaitk_num_streams.restype = i32
    L.gaitk_stream_in_dim.argtypes = [vp, i32]; L.gaitk_stream_in_dim.restype = i32
    L.gaitk_stream_in_len.argtypes = [vp, i32]; L.gaitk_stream_in_len.restype = i32
    L.gaitk_stream_geometry.argtypes = [vp, i32, i32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(sz)]
    L.gaitk_stream_geometry.restype = i32
    L.gaitk_workspace_bytes.argtypes = [vp, i32]; L.gaitk_workspace_bytes.restype = sz
    L.gaitk_forward.argtypes = [vp, vp, pp, pp, i32, u32, pp, i32, vp]; L.gaitk_forward.restype = i32
    L.gaitk_loss.argtypes = [vp, vp, i32, i32, C.POINTER(LossDesc), vp, vp, vp, vp, vp]; L.gaitk_loss.restype = i32
    L.gaitk_backward.argtypes = [vp, vp, pp, pp, i32, u32, pp, vp, vp, sz, i32, vp]; L.gaitk_backward.restype = i32
    L.gaitk_step_grads.argtypes = [vp, vp, pp, pp, pp, i32, C.POINTER(LossDesc), pp, vp, u32, u32, f32, f32, pp, vp,
                                   vp, sz, i32, vp]
    L.gaitk_step_grads.restype = i32
    L.gaitk_loss_denominators.argtypes = [pp, C.POINTER(C.c_int), i32, C.POINTER(LossDesc), vp, vp]
    L.gaitk_loss_denominators.restype = i32
    L.gaitk_step_update.argtypes = [vp, vp, vp, vp, u32, f32, f32, f32, f32, f32, vp, vp, i32, vp]
    L.gaitk_step_update.restype = i32
    L.gaitk_p2p_allreduce.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]; L.gaitk_p2p_allreduce.restype = i32
    L.gaitk_cagrad.argtypes = [vp, i32, i32, f32, f32, vp, vp, i32, vp]; L.gaitk_cagrad.restype = i32
    L.gaitk_cagrad_solve_host.argtypes = [C.POINTER(C.c_float), i32, f32, i32, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.gaitk_cagrad_solve_host.restype = i32
    L.gaitk_sgd.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_uint8), f32, f32, f32, vp]; L.gaitk_sgd.restype = i32
    L.gaitk_window_indices.argtypes = [i64, i64, i64, C.POINTER(i64), i64]; L.gaitk_window_indices.restype = i64
    L.gaitk_stats_accumulate.argtypes = [vp, i64, i32, vp, vp]; L.gaitk_stats_accumulate.restype = i32
    L.gaitk_stats_finalize.argtypes = [vp, i32, vp, vp, vp]; L.gaitk_stats_finalize.restype = i32
    L.gaitk_normalize_frames.argtypes = [vp, i64, i32, vp, vp, vp, vp]; L.gaitk_normalize_frames.restype = i32
    L.gaitk_window_gather.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]; L.gaitk_window_gather.restype = i32
    L.gaitk_mask_eval.argtypes = [pp, pp, i32, i32, vp, vp]; L.gaitk_mask_eval.restype = i32
    L.gaitk_fog_prepare_pose.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]; L.gaitk_fog_prepare_pose.restype = i32
    L.gaitk_fog_prepare_sensor.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]; L.gaitk_fog_prepare_sensor.restype = i32
    L.gaitk_umma_selftest.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, vp]; L.gaitk_umma_selftest.restype = i32
    L.gaitk_umma_bench.argtypes = [vp, i32, i32, i32, i32, vp, vp]; L.gaitk_umma_bench.restype = i32
    L.gaitk_umma_bench_multi.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]; L.gaitk_umma_bench_multi.restype = i32
    L.gaitk_umma_selftest_bf16.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, vp]; L.gaitk_umma_selftest_bf16.restype = i32
    L.gaitk_stage_create.argtypes = [C.POINTER(StageDesc), i32, pp]; L.gaitk_stage_create.restype = i32
    L.gaitk_stage_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]; L.gaitk_stage_forward.restype = i32
    L.gaitk_stage_backward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, sz, vp]; L.gaitk_stage_backward.restype = i32
    L.gaitk_xattn_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp]; L.gaitk_xattn_forward.restype = i32
    L.gaitk_xattn_backward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]; L.gaitk_xattn_backward.restype = i32
    L.gaitk_linear_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]; L.gaitk_linear_forward.restype = i32
    L.gaitk_linear_workspace_bytes.argtypes = [i32, i32, i32]; L.gaitk_linear_workspace_bytes.restype = sz
    L.gaitk_linear_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]; L.gaitk_linear_backward.restype = i32
    L.gaitk_adam.argtypes = [pp, pp, pp, pp, C.POINTER(i64), i32, f32, f32, f32, f32, f32, i32, vp]; L.gaitk_adam.restype = i32
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().gaitk_last_error().decode("utf-8", "replace")
        kind = {-1: "GAITK_E_BADARG", -2: "GAITK_E_SHAPE", -3: "GAITK_E_DTYPE", -4: "GAITK_E_ARCH",
                -5: "GAITK_E_STATE"}.get(rc, f"cudaError {rc}" if rc > 0 else str(rc))
        raise GaitkError(f"{what or 'gaitk'} failed: {kind}: {msg}")


def ptr_array(ptrs):
    """list of ints (device pointers, 0 = NULL) -> void*[]"""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def stream_handle():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
