"""ctypes binding of libgsr_b200.so — the C ABI declared in include/gsr.h.

There is deliberately no fallback: if the shared library is missing the import raises, and compute
entry points raise RuntimeError with gsr_last_error() on any failure.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsr_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)

GSR_OK = 0
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class FwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
        ("scale_modifier", C.c_float), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("cam_pos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("prefiltered", C.c_int),
        ("out_color", C.c_void_p), ("out_depth", C.c_void_p), ("out_acc", C.c_void_p),
        ("radii", C.c_void_p), ("debug", C.c_int), ("activations", C.c_int),
    ]


class BwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("R", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("scales", C.c_void_p), ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p), ("cam_pos", C.c_void_p), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", C.c_void_p), ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p),
        ("image_buffer", C.c_void_p), ("dL_dpix", C.c_void_p), ("dL_dacc", C.c_void_p),
        ("dL_dmean2D", C.c_void_p), ("dL_dconic", C.c_void_p), ("dL_dopacity", C.c_void_p),
        ("dL_dcolor", C.c_void_p), ("dL_dmean3D", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("dL_dsh", C.c_void_p), ("dL_dscale", C.c_void_p), ("dL_drot", C.c_void_p), ("debug", C.c_int),
        ("flags", C.c_int), ("row_state", C.c_void_p), ("activations", C.c_int), ("opacities", C.c_void_p), ("range_first", C.c_int),
        ("range_count", C.c_int),
    ]


BWD_OUTPUTS_UNINITIALIZED = 1
BWD_ACCUMULATORS_ZEROED = 2
BWD_SH_FACTORED = 4
BWD_ONLY_BLEND = 8
BWD_ONLY_GEOM = 16
ACT_OPACITY_SIGMOID, ACT_SCALE_EXP, ACT_ROTATION_NORMALIZE = 1, 2, 4
ACT_ALL = 7
WINDOW_CH_DONE, WINDOW_CH_PUSHED = 2, 3
ABI_VERSION = 5  # include/gsr.h GSR_ABI_VERSION this binding was written against


class WindowRows(C.Structure):
    _fields_ = [("union_words", C.c_void_p), ("first_float", C.c_size_t), ("n_rows", C.c_uint32),
                ("floats_per_row", C.c_uint32)]


class WindowComm(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("pads", C.c_void_p * 16)]


class AdamGroup(C.Structure):
    _fields_ = [("first", C.c_size_t), ("count", C.c_size_t), ("lr", C.c_float), ("period", C.c_uint), ("phase_lo", C.c_uint),
                ("phase_hi", C.c_uint), ("lr_other", C.c_float)]


class VgprParams(C.Structure):
    _fields_ = [("grid", C.c_float), ("num_gp_side", C.c_int), ("neighbour", C.c_int), ("kernel_size", C.c_float),
                ("variance_sensor", C.c_float), ("scale_factor", C.c_float), ("full_cover", C.c_int)]


class ArrayView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride", C.c_size_t), ("count", C.c_size_t)]


class GeomView(C.Structure):
    _fields_ = [(n, ArrayView) for n in (
        "means2D", "extent", "conic_opacity", "rgb", "depths", "clamped", "internal_radii", "tiles_touched",
        "point_offsets")]


class BinningView(C.Structure):
    _fields_ = [(n, ArrayView) for n in (
        "point_list_keys_unsorted", "point_list_unsorted", "point_list_keys", "point_list", "instance_records",
        "block_candidates")]


class ImageView(C.Structure):
    _fields_ = [(n, ArrayView) for n in ("accum_alpha", "n_contrib", "ranges", "tile_last", "tile_order", "block_count",
                                         "last_group", "summary")]


class AsyncStatus(C.Structure):
    _fields_ = [("num_rendered", C.c_uint32), ("longest_list", C.c_uint32), ("long_lists", C.c_uint32),
                ("overflow", C.c_uint32)]


# every symbol include/gsr.h declares: (name, restype, argtypes)
SYMBOLS = {
    "gsr_abi_version": (C.c_int, []),
    "gsr_last_error": (C.c_char_p, []),
    "gsr_sizeof": (C.c_size_t, [C.c_int]),
    "gsr_required_geom": (C.c_size_t, [C.c_int]),
    "gsr_required_image": (C.c_size_t, [C.c_int, C.c_int]),
    "gsr_required_binning": (C.c_size_t, [C.c_int]),
    "gsr_forward": (C.c_int, [C.POINTER(FwdArgs), ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p,
                              C.c_void_p, c_int_p]),
    "gsr_forward_presized": (C.c_int, [C.POINTER(FwdArgs), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "gsr_backward": (C.c_int, [C.POINTER(BwdArgs), C.c_void_p]),
    "gsr_backward_prepare": (C.c_int, [C.POINTER(BwdArgs), C.c_void_p]),
    "gsr_window_allreduce_multimem": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "gsr_window_allreduce_p2p": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gsr_window_visibility_words": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gsr_window_union_words": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsr_window_allreduce_multimem_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsr_window_allreduce_p2p_rows": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p]),
    "gsr_window_slot_floats": (C.c_size_t, [C.c_int]),
    "gsr_window_pad_words": (C.c_size_t, []),
    "gsr_window_status": (C.c_int, [C.POINTER(WindowComm), C.POINTER(C.c_uint32)]),
    "gsr_window_wait": (C.c_int, [C.POINTER(WindowComm), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gsr_window_push_rgb": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                      C.POINTER(WindowComm), C.c_void_p]),
    "gsr_window_expand_sh": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsr_window_allreduce_rest": (C.c_int, [C.POINTER(WindowComm), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.c_int,
                                            C.POINTER(C.c_void_p), C.c_void_p]),
    "gsr_image_loss_scratch_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "gsr_image_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "gsr_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AdamGroup), C.c_int, C.c_float,
                                C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "gsr_vgpr_gaussians": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VgprParams), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsr_mark_visible": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsr_view_geom": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GeomView)]),
    "gsr_view_binning": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(BinningView)]),
    "gsr_view_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(ImageView)]),
    "gsr_configure": (C.c_int, [C.c_char_p, C.c_int]),
    "gsr_config_get": (C.c_int, [C.c_char_p, c_int_p]),
    "gsr_stage_timing_enable": (None, [C.c_int]),
    "gsr_stage_times": (C.c_int, [c_float_p, C.c_int]),
    "gsr_shim_set_stream": (None, [C.c_void_p]),
    "gsr_shim_get_stream": (C.c_void_p, []),
}

STAGES = ("preprocess", "scan", "duplicate", "sort", "finalize", "blend_fwd", "blend_bwd", "geom_bwd")

_lib = None


def load() -> C.CDLL:
    """Load libgsr_b200.so (built by gs-livm_b200/build.py). Raises if it is missing — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first (python gs-livm_b200/build.py or "
                "__graft_entry__.build()). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def configure(**knobs: int) -> None:
    """gsr_configure for every keyword (include/gsr.h lists the knobs); a negative value restores the default."""
    lib = load()
    for name, value in knobs.items():
        check(lib.gsr_configure(name.encode(), int(value)), "gsr_configure(%s)" % name)


def check(status: int, where: str) -> None:
    if status != GSR_OK:
        msg = load().gsr_last_error()
        raise RuntimeError(f"{where} failed ({status}): {msg.decode() if msg else '?'}")
