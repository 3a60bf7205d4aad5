"""ctypes binding of libd2r_b200.so (the C ABI declared in include/d2r_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised (the reference's pybind11 layer turns std::runtime_error into RuntimeError too).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libd2r_b200.so")

EXPORTS = [
    "d2r_model_load", "d2r_model_free", "d2r_model_set_min_transmittance", "d2r_model_get_bitfield", "d2r_model_get_occupied_aabb",
    "d2r_view_prepare", "d2r_view_free", "d2r_view_get_dirs",
    "d2r_render", "d2r_render_ex", "d2r_render_composite", "d2r_render_composite_ex",
    "d2r_clip_preprocess", "d2r_clip_preprocess_delta", "d2r_clip_load", "d2r_clip_free", "d2r_clip_encode", "d2r_score", "d2r_phys_check", "d2r_gemm_f16", "d2r_profile_enable", "d2r_profile_read", "d2r_profile_read_stats",
    "d2r_last_error", "d2r_launch_count", "d2r_version",
]


class ModelCfg(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features_per_level", C.c_int32), ("log2_hashmap_size", C.c_int32),
                ("base_resolution", C.c_int32), ("per_level_scale", C.c_float), ("max_cascade", C.c_int32),
                ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
                ("render_aabb_min", C.c_float * 3), ("render_aabb_max", C.c_float * 3),
                ("render_aabb_to_local", C.c_float * 9), ("cone_angle_constant", C.c_float),
                ("min_transmittance", C.c_float), ("depth_scale", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("focal", C.c_float * 2), ("screen_center", C.c_float * 2),
                ("lens_mode", C.c_int32), ("lens_params", C.c_float * 4)]


class ClipCfg(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("patch_size", C.c_int32), ("hidden", C.c_int32), ("heads", C.c_int32),
                ("layers", C.c_int32), ("mlp", C.c_int32), ("proj", C.c_int32), ("ln_eps", C.c_float),
                ("max_batch", C.c_int32)]


class PhysCfg(C.Structure):
    _fields_ = [("dataset_scale", C.c_float), ("dataset_offset", C.c_float * 3), ("scene_centre_z", C.c_float),
                ("unsup_thresh", C.c_float), ("p_dist", C.c_float), ("stability_check", C.c_int32)]


_lib = None


def lib():
    """Load the shared library (once).  Raises if it is not built: the product never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python dream2real_b200/csrc/build.py` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, f4 = C.c_void_p, C.c_int, C.POINTER(C.c_float)
    sigs = {
        "d2r_last_error": (C.c_char_p, []),
        "d2r_version": (C.c_char_p, []),
        "d2r_launch_count": (C.c_ulonglong, [i]),
        "d2r_model_load": (i, [vp, C.c_size_t, vp, C.c_size_t, C.POINTER(ModelCfg), i, C.POINTER(vp)]),
        "d2r_model_free": (None, [vp]),
        "d2r_model_set_min_transmittance": (i, [vp, C.c_float]),
        "d2r_model_get_bitfield": (i, [vp, vp, C.c_size_t]),
        "d2r_model_get_occupied_aabb": (i, [vp, vp]),
        "d2r_view_prepare": (i, [C.POINTER(Camera), i, C.POINTER(vp)]),
        "d2r_view_free": (None, [vp]),
        "d2r_view_get_dirs": (i, [vp, vp]),
        "d2r_render": (i, [vp, vp, vp, i, f4, vp, vp, vp, vp]),
        "d2r_render_ex": (i, [vp, vp, vp, i, f4, vp, vp, vp, vp, vp]),
        "d2r_render_composite": (i, [vp, vp, vp, i, f4, vp, vp, vp, vp, vp]),
        "d2r_render_composite_ex": (i, [vp, vp, vp, i, f4, vp, vp, vp, vp, vp, vp, vp]),
        "d2r_clip_preprocess": (i, [vp, i, i, i, i, i, i, f4, f4, vp, vp, vp]),
        "d2r_clip_preprocess_delta": (i, [vp, i, i, i, i, i, i, f4, f4, vp, vp, vp, vp]),
        "d2r_clip_load": (i, [C.POINTER(ClipCfg), C.POINTER(vp), i, i, C.POINTER(vp)]),
        "d2r_clip_free": (None, [vp]),
        "d2r_clip_encode": (i, [vp, vp, i, vp, vp]),
        "d2r_score": (i, [vp, vp, i, i, i, C.c_float, i, vp, vp, vp]),
        "d2r_phys_check": (i, [vp, vp, i, vp, vp, vp, i, C.POINTER(PhysCfg), vp, vp]),
        "d2r_profile_enable": (i, [i, i]),
        "d2r_profile_read": (i, [i, vp, vp, vp, vp]),
        "d2r_profile_read_stats": (i, [i, vp]),
        "d2r_gemm_f16": (i, [vp, i, vp, i, i, i, i, vp, i, vp, i, vp]),
    }
    assert set(sigs) == set(EXPORTS)
    missing = []
    for name, (res, args) in sigs.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export {missing}: stale build? run dream2real_b200/csrc/build.py --force")
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().d2r_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'libd2r_b200'} failed ({rc}): {msg}")


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def f4(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def stream_ptr(stream: Optional[object] = None):
    """cudaStream_t of torch's current stream (or the given torch.cuda.Stream) as void*."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def launch_count(reset: bool = False) -> int:
    return int(lib().d2r_launch_count(1 if reset else 0))
