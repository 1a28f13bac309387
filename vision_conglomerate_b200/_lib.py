"""ctypes binding of ``csrc/libboxgeom.so`` (the C ABI declared in ``include/boxgeom.h``).

There is no CPU or PyTorch fallback: if the shared library is missing this module raises, and every
operator in :mod:`vision_conglomerate_b200.ops` refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libboxgeom.so")

BG_MAX_ANCHORS = 8
BG_MAX_TRACKED = 64
STATUS_GROUP_RANGE = 1
STATUS_MASK_SPACE = 2
STATUS_NEED_GENERAL = 4

# every symbol include/boxgeom.h declares (tests check that the library exports each of them)
SYMBOLS = (
    "bg_strerror", "bg_version", "bg_launch_count", "bg_sizeof_detect_params", "bg_sizeof_loss_params", "bg_profile_events", "bg_profile_events_loss",
    "bg_profile_stamps", "bg_profile_stamps_per_image", "bg_profile_decode_cycles",
    "bg_batched_nms_workspace_bytes", "bg_batched_nms",
    "bg_host_mapped_ptr", "bg_detect_workspace_bytes", "bg_detect", "bg_post_process", "bg_decode_scale", "bg_decode_scale_ex", "bg_decode_rows", "bg_bbox_to_size", "bg_decode_train_bwd", "bg_decode_train_bwd_ex",
    "bg_assign_workspace_bytes", "bg_assign_targets", "bg_assign_ex_workspace_bytes", "bg_assign_targets_ex",
    "bg_ciou_fwd", "bg_ciou_bwd",
    "bg_loss_workspace_bytes", "bg_loss_fwd", "bg_loss_bwd", "bg_loss_clear_grads", "bg_loss_pack", "bg_loss_combine",
    "bg_ratio_metrics",
    "bg_seg_loss_workspace_bytes", "bg_seg_loss_fwd", "bg_seg_loss_bwd", "bg_sizeof_seg_params", "bg_seg_masks",
)


class DetectParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("na", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32),
        ("og_H", C.c_int32), ("og_W", C.c_int32),
        ("ny", C.c_int32 * 3), ("nx", C.c_int32 * 3),
        ("anchors", ((C.c_float * 2) * BG_MAX_ANCHORS) * 3),
        ("box_allowance", C.c_float),
        ("score_threshold", C.c_float),
        ("iou_threshold", C.c_double),
        ("n_tracked", C.c_int32),
        ("tracked", C.c_int32 * BG_MAX_TRACKED),
        ("order", C.c_int32),
        ("variant", C.c_int32),
        ("nms_path", C.c_int32),
        ("extra_cols", C.c_int32),
        ("throughput", C.c_int32),
        ("nms_stream", C.c_void_p),
        ("nms_event", C.c_void_p),
        ("host_flag", C.c_void_p),
        ("host_flag_value", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class LossParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("na", C.c_int32),
        ("ny", C.c_int32 * 3), ("nx", C.c_int32 * 3),
        ("anchors", ((C.c_float * 2) * BG_MAX_ANCHORS) * 3),
        ("anchor_t", C.c_float), ("edge_t", C.c_float), ("label_smoothing", C.c_float),
        ("box_w", C.c_float), ("conf_w", C.c_float), ("class_w", C.c_float),
        ("scale_w", C.c_float * 3),
        ("nt", C.c_int64),
        ("input_form", C.c_int32),
        ("extra_cols", C.c_int32),
    ]


class SegParams(C.Structure):
    """bg_seg_params: the mask term of SegmentationLoss on top of the fused detection loss."""
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("na", C.c_int32), ("K", C.c_int32),
        ("extra_cols", C.c_int32),
        ("ny", C.c_int32 * 3), ("nx", C.c_int32 * 3),
        ("anchors", ((C.c_float * 2) * BG_MAX_ANCHORS) * 3),
        ("anchor_t", C.c_float), ("edge_t", C.c_float),
        ("Hp", C.c_int32), ("Wp", C.c_int32), ("Hm", C.c_int32), ("Wm", C.c_int32),
        ("scale_w", C.c_float * 3),
        ("seg_w", C.c_float),
        ("nt", C.c_int64),
    ]


Ptr3 = C.c_void_p * 3

LOSS_DECODED, LOSS_RAW, LOSS_RAW_SPLIT = 0, 1, 2
LOSS_BWD_PRECLEARED = 1


class HeadPtrs(C.Structure):
    """One scale of the loss inputs / gradients (bg_head_ptrs, bg_head_grads): interleaved forms use `obj` only."""
    _fields_ = [("obj", C.c_void_p), ("cls", C.c_void_p), ("box", C.c_void_p)]


HeadPtrs3 = HeadPtrs * 3


def build(force: bool = False) -> str:
    """Compile libboxgeom.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "boxgeom.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-B", "libboxgeom.so"])
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for the box-geometry path)")
    L = C.CDLL(SO_PATH)
    vp, sz, i64, i32, f32, f64 = C.c_void_p, C.c_size_t, C.c_int64, C.c_int32, C.c_float, C.c_double
    L.bg_strerror.argtypes = [C.c_int]
    L.bg_strerror.restype = C.c_char_p
    L.bg_version.restype = C.c_int
    L.bg_launch_count.restype = C.c_uint64
    L.bg_sizeof_detect_params.restype = sz
    L.bg_sizeof_loss_params.restype = sz
    if L.bg_sizeof_detect_params() != C.sizeof(DetectParams) or L.bg_sizeof_loss_params() != C.sizeof(LossParams):
        raise RuntimeError("libboxgeom.so: parameter struct layout differs from the ctypes binding (stale build?)")
    L.bg_profile_events.argtypes = [vp, vp]
    L.bg_profile_events.restype = None
    L.bg_profile_events_loss.argtypes = [vp, vp]
    L.bg_profile_events_loss.restype = None
    L.bg_profile_stamps.argtypes = [vp]
    L.bg_profile_stamps.restype = None
    L.bg_profile_stamps_per_image.restype = C.c_int
    L.bg_profile_decode_cycles.argtypes = [vp]
    L.bg_profile_decode_cycles.restype = None
    L.bg_batched_nms_workspace_bytes.argtypes = [i64, i64, sz]
    L.bg_batched_nms_workspace_bytes.restype = sz
    L.bg_batched_nms.argtypes = [vp, vp, vp, i64, f64, i64, vp, vp, vp, sz, sz, vp]
    L.bg_host_mapped_ptr.argtypes = [vp]
    L.bg_host_mapped_ptr.restype = vp
    L.bg_detect_workspace_bytes.argtypes = [C.POINTER(DetectParams), sz]
    L.bg_detect_workspace_bytes.restype = sz
    L.bg_detect.argtypes = [vp, vp, vp, C.POINTER(DetectParams), vp, vp, vp, vp, vp, sz, sz, vp]
    L.bg_post_process.argtypes = [vp, C.POINTER(DetectParams), vp, vp, vp, vp, vp, sz, sz, vp]
    L.bg_decode_scale.argtypes = [vp, vp, i32, i32, i32, i32, i32, C.POINTER(f32), i32, i32, i32, i32, i32, vp]
    L.bg_decode_scale_ex.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(f32), i32, i32, i32, i32, i32, vp]
    L.bg_decode_train_bwd.argtypes = [vp, vp, vp, i64, i32, vp]
    L.bg_decode_train_bwd_ex.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp]
    L.bg_bbox_to_size.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.bg_decode_rows.argtypes = [vp, vp, vp, C.POINTER(DetectParams), vp, i64, vp, vp]
    L.bg_assign_workspace_bytes.argtypes = [i64, i32]
    L.bg_assign_workspace_bytes.restype = sz
    L.bg_assign_targets.argtypes = [vp, i64, i32, i32, C.POINTER(f32), i32, f32, f32, vp, vp, vp, vp, i64, vp, vp, sz, vp]
    L.bg_assign_ex_workspace_bytes.argtypes = [i64, i32, i32]
    L.bg_assign_ex_workspace_bytes.restype = sz
    L.bg_assign_targets_ex.argtypes = [vp, i64, i32, i32, i32, C.POINTER(f32), i32, f32, f32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp,
                                       vp, sz, vp]
    L.bg_ciou_fwd.argtypes = [vp, vp, i64, f32, vp, vp]
    L.bg_ciou_bwd.argtypes = [vp, vp, vp, i64, f32, vp, vp]
    L.bg_loss_workspace_bytes.argtypes = [C.POINTER(LossParams)]
    L.bg_loss_workspace_bytes.restype = sz
    L.bg_loss_fwd.argtypes = [C.POINTER(HeadPtrs), vp, C.POINTER(LossParams), vp, vp, vp, vp, vp, sz, vp]
    L.bg_loss_bwd.argtypes = [C.POINTER(HeadPtrs), C.POINTER(LossParams), vp, f32, C.POINTER(HeadPtrs), i32, vp, sz, vp]
    L.bg_loss_clear_grads.argtypes = [C.POINTER(LossParams), C.POINTER(HeadPtrs), vp]
    L.bg_loss_pack.argtypes = [vp, C.POINTER(i64), i32, vp, vp]
    L.bg_loss_combine.argtypes = [vp, C.POINTER(LossParams), vp, vp]
    L.bg_ratio_metrics.argtypes = [vp, i64, C.POINTER(f32), i32, f32, vp, vp]
    L.bg_sizeof_seg_params.restype = sz
    if L.bg_sizeof_seg_params() != C.sizeof(SegParams):
        raise RuntimeError("libboxgeom.so: bg_seg_params layout differs from the ctypes binding (stale build?)")
    L.bg_seg_loss_workspace_bytes.argtypes = [C.POINTER(SegParams)]
    L.bg_seg_loss_workspace_bytes.restype = sz
    L.bg_seg_loss_fwd.argtypes = [C.POINTER(vp), vp, vp, vp, C.POINTER(SegParams), vp, vp, vp, vp, sz, vp]
    L.bg_seg_loss_bwd.argtypes = [C.POINTER(vp), vp, vp, C.POINTER(SegParams), vp, C.POINTER(vp), vp, vp, sz, vp]
    L.bg_seg_masks.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, i32, i32, vp, vp, vp]
    for name in ("bg_batched_nms", "bg_detect", "bg_post_process", "bg_decode_scale", "bg_decode_scale_ex", "bg_decode_rows", "bg_bbox_to_size", "bg_decode_train_bwd", "bg_decode_train_bwd_ex", "bg_assign_targets", "bg_assign_targets_ex", "bg_ciou_fwd", "bg_ciou_bwd",
                 "bg_loss_fwd", "bg_loss_bwd", "bg_loss_clear_grads", "bg_loss_pack", "bg_loss_combine", "bg_ratio_metrics",
                 "bg_seg_loss_fwd", "bg_seg_loss_bwd", "bg_seg_masks"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().bg_strerror(rc).decode()} (code {rc})")


def launch_count() -> int:
    return int(lib().bg_launch_count())
