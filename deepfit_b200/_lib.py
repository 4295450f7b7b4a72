"""ctypes binding of ``libdeepfit_b200.so`` (the C ABI declared in ``include/deepfit_b200.h``).

There is deliberately no fallback: if the shared library cannot be found or built, importing any
compute entry point raises.  PyTorch is used by the callers only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

DFB_NUM_LAYERS = 20
LAYER_NAMES = [
    "QSTN_CONV1", "QSTN_CONV2", "QSTN_CONV3", "QSTN_FC1", "QSTN_FC2", "QSTN_FC3",
    "PF_CONV1", "PF_CONV2",
    "STN_CONV1", "STN_CONV2", "STN_CONV3", "STN_FC1", "STN_FC2", "STN_FC3",
    "ENC_CONV2", "ENC_CONV3",
    "HEAD_CONV1", "HEAD_CONV2", "HEAD_CONV3", "HEAD_CONV4",
]
assert len(LAYER_NAMES) == DFB_NUM_LAYERS

WEIGHT_MODE = {"sigmoid": 0, "tanh": 1}
PRECISION = {"bf16x3": 0, "bf16": 1, "f16f8": 2, "f16x3": 3, "f16mix": 4}
# DFB_MIX_* bits of include/deepfit_b200.h: the chain layers that "f16mix" runs with the hi*hi product only
MIX_BITS = {"qstn_small": 1, "qstn_max": 2, "stn_small": 4, "stn_max": 8, "enc_small": 16, "enc_max": 32}


def precision_code(precision: str) -> int:
    """'f16x3' ... or 'f16mix' / 'f16mix:qstn_max+stn_max' (an explicit DFB_MIX_* mask, for the precision ladder)."""
    name, _, layers = precision.partition(":")
    if name not in PRECISION or (layers and name != "f16mix"):
        raise ValueError("precision must be one of %s (f16mix optionally as 'f16mix:<layer>+<layer>' with layers %s)"
                         % (sorted(PRECISION), sorted(MIX_BITS)))
    mask = 0
    for layer in filter(None, layers.split("+")):
        if layer not in MIX_BITS:
            raise ValueError("unknown f16mix layer %r; known: %s" % (layer, sorted(MIX_BITS)))
        mask |= MIX_BITS[layer]
    return PRECISION[name] | (mask << 8)

EXPORTS = [
    "dfb_abi_version", "dfb_last_error_string", "dfb_device_check",
    "dfb_grid_create", "dfb_grid_update", "dfb_grid_destroy", "dfb_grid_morton_order", "dfb_knn", "dfb_ball_count", "dfb_ball_fill",
    "dfb_patch_pca", "dfb_patch_pca_ragged",
    "dfb_model_create", "dfb_model_destroy", "dfb_model_forward", "dfb_model_status",
    "dfb_model_profile", "dfb_model_profile_read", "dfb_launch_count",
    "dfb_wls_fit", "dfb_principal_curvatures", "dfb_neighbor_normals",
    "dfb_pipeline_create", "dfb_pipeline_run", "dfb_pipeline_destroy",
    "dfb_savetxt_f32", "dfb_savetxt_f64", "dfb_loadtxt_f32",
]


class DfbLayer(C.Structure):
    _fields_ = [("h_weight", C.c_void_p), ("h_bias", C.c_void_p),
                ("in_features", C.c_int32), ("out_features", C.c_int32)]


class DfbModelDesc(C.Structure):
    _fields_ = [("k", C.c_int32), ("weight_mode", C.c_int32), ("precision", C.c_int32),
                ("max_batch", C.c_int32), ("layers", DfbLayer * DFB_NUM_LAYERS)]


class DfbPipelineOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_normals", "d_curv", "d_beta", "d_weights", "d_trans", "d_trans2", "d_U", "d_rad",
                                          "d_dirs_world", "d_info")]


DFB_PIPE_KEEP_QSTN = 1
DFB_PIPE_FILE_ORDER = 2
ABI_VERSION = 2


class DeepFitB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("deepfit_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get("DEEPFIT_B200_LIB")   # A/B experiments with another build of the same ABI
    if override:
        path = override
    elif not os.path.isfile(path) or (not _build.is_current() and os.environ.get("DEEPFIT_B200_NO_REBUILD") != "1"):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001 - re-raised with context below when no library exists
            if not os.path.isfile(path):
                raise ImportError(
                    "deepfit_b200: CUDA library %s is missing and could not be built (%s). "
                    "There is no CPU fallback; run `python -m deepfit_b200.build`." % (path, e)) from e
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.dfb_abi_version.restype = C.c_int
    lib.dfb_last_error_string.restype = C.c_char_p
    lib.dfb_device_check.restype = C.c_int
    lib.dfb_grid_create.argtypes = [vp, i64, vp, C.POINTER(vp)]
    lib.dfb_grid_destroy.argtypes = [vp]
    lib.dfb_grid_update.argtypes = [vp, vp, vp]
    lib.dfb_grid_morton_order.argtypes = [vp, i64, i64, vp, vp]
    lib.dfb_knn.argtypes = [vp, vp, i64, i64, i32, vp, vp, vp, vp]
    lib.dfb_ball_count.argtypes = [vp, vp, i64, i64, vp, C.c_double, vp, vp]
    lib.dfb_ball_fill.argtypes = [vp, vp, i64, i64, vp, C.c_double, vp, vp, vp]
    lib.dfb_patch_pca.argtypes = [vp, i64, vp, vp, i64, i64, i32, vp, i32, vp, vp, vp]
    lib.dfb_patch_pca_ragged.argtypes = [vp, i64, vp, vp, vp, i64, i64, i32, vp, i32, vp, vp, vp]
    lib.dfb_model_create.argtypes = [C.POINTER(DfbModelDesc), vp, C.POINTER(vp)]
    lib.dfb_model_destroy.argtypes = [vp]
    lib.dfb_model_forward.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    lib.dfb_model_status.argtypes = [vp, vp]
    lib.dfb_model_profile.argtypes = [vp, i32]
    lib.dfb_model_profile_read.argtypes = [vp, vp, vp, i32, vp]
    lib.dfb_launch_count.argtypes = [i32]
    lib.dfb_dbg_set.argtypes = [C.c_int, C.c_int]
    lib.dfb_dbg_gemm.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.dfb_dbg_gemm.restype = C.c_int
    lib.dfb_dbg_model_dump.argtypes = [vp, i32, i64, i32, vp, vp]
    lib.dfb_dbg_model_dump.restype = C.c_int
    lib.dfb_wls_fit.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.dfb_principal_curvatures.argtypes = [vp, i64, i32, vp, vp, vp]
    lib.dfb_pipeline_create.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(vp)]
    lib.dfb_pipeline_run.argtypes = [vp, i64, i64, C.POINTER(DfbPipelineOutputs), vp]
    lib.dfb_pipeline_destroy.argtypes = [vp]
    lib.dfb_neighbor_normals.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    lib.dfb_savetxt_f32.argtypes = [C.c_char_p, vp, i64, i32]
    lib.dfb_savetxt_f64.argtypes = [C.c_char_p, vp, i64, i32]
    lib.dfb_loadtxt_f32.argtypes = [C.c_char_p, vp, i64, C.POINTER(i64), C.POINTER(i32)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name == "dfb_launch_count":
            fn.restype = C.c_int64
        elif name != "dfb_last_error_string":
            fn.restype = C.c_int
    if lib.dfb_abi_version() != ABI_VERSION:
        raise ImportError("deepfit_b200: ABI version mismatch in %s" % path)
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().dfb_last_error_string()
        raise DeepFitB200Error(rc, msg.decode("utf-8", "replace") if msg else "")


def require_device() -> None:
    """Raise unless an sm_100 device is current -- the product has no CPU path."""
    check(load().dfb_device_check())
