"""ctypes binding of include/vkgsb.h.  There is no fallback: a missing or unloadable libvkgsb.so is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VKGSB_LIB selects another build of the same library (A/B kernel experiments, tools/build_variant.sh); still no fallback
LIB_PATH = os.environ.get("VKGSB_LIB") or os.path.join(_HERE, "lib", "libvkgsb.so")

OK, ERR_INVALID, ERR_CUDA, ERR_IO, ERR_CAPACITY, ERR_NO_SCENE, ERR_CANCELLED = range(7)
BLEND_FP32, BLEND_UNORM8 = 0, 1
EXTERNAL_OPAQUE_FD, EXTERNAL_CUDA_POSIX_FD = 0, 1
FORMAT_RGBA8, FORMAT_BGRA8 = 0, 1
OPT_STAGE_TIMING, OPT_BLEND_MODE, OPT_PIXEL_FORMAT, OPT_BAND_Y0, OPT_BAND_Y1, OPT_KEEP_INSTANCES, OPT_BAND_CULL, OPT_COUNT_FRAGMENTS, OPT_UNORM8_CUT_EXP, OPT_L2_PIN_MB, OPT_SPATIAL_ORDER = range(11)


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_splats", C.c_uint32),
                ("max_width", C.c_uint32), ("max_height", C.c_uint32), ("max_pairs", C.c_uint64)]


class CameraBlock(C.Structure):
    """vkgsb_camera: shader::Camera (uniforms.h:10-15) + the model push constant."""
    _fields_ = [("projection", C.c_float * 16), ("view", C.c_float * 16), ("camera_position", C.c_float * 3),
                ("pad0", C.c_float), ("model", C.c_float * 16)]


class Stats(C.Structure):
    _fields_ = [("total_point_count", C.c_uint32), ("loaded_point_count", C.c_uint32),
                ("visible_point_count", C.c_uint32), ("pair_overflow", C.c_uint32), ("pair_count", C.c_uint64),
                ("ms_project", C.c_float), ("ms_sort", C.c_float), ("ms_bin", C.c_float), ("ms_blend", C.c_float),
                ("ms_total", C.c_float), ("frame_counter", C.c_uint64), ("blend_retries", C.c_uint32),
                ("pad0", C.c_uint32), ("fragment_count", C.c_uint64), ("ms_cull", C.c_float),
                ("pad1", C.c_uint32)]


# every entry point include/vkgsb.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "vkgsb_last_error": (C.c_char_p, []),
    "vkgsb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vkgsb_create": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(_P)]),
    "vkgsb_create_ex": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "vkgsb_destroy": (None, [_P]),
    "vkgsb_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "vkgsb_load_ply": (C.c_int, [_P, C.c_char_p]),
    "vkgsb_load_ply_async": (C.c_int, [_P, C.c_char_p]),
    "vkgsb_load_progress": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    "vkgsb_cancel_load": (C.c_int, [_P]),
    "vkgsb_wait_load": (C.c_int, [_P]),
    "vkgsb_upload_splats": (C.c_int, [_P, C.c_uint32, _P, _P]),
    "vkgsb_set_camera": (C.c_int, [_P, C.POINTER(CameraBlock)]),
    "vkgsb_set_viewport": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "vkgsb_set_lines": (C.c_int, [_P, C.c_uint32, _P, _P, _P]),
    "vkgsb_draw": (C.c_int, [_P, _P, C.c_int, _P]),
    "vkgsb_draw_batch": (C.c_int, [_P, C.c_uint32, C.POINTER(CameraBlock), _P, C.c_size_t, C.c_int, _P]),
    "vkgsb_image_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "vkgsb_sync": (C.c_int, [_P]),
    "vkgsb_wait_frame": (C.c_int, [_P, C.c_uint32]),
    "vkgsb_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "vkgsb_row_histogram": (C.c_int, [_P, _P, C.c_uint32]),
    "vkgsb_read_sorted": (C.c_int, [_P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "vkgsb_read_instances": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "vkgsb_read_scene": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "vkgsb_read_order": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "vkgsb_sort_storage_bytes": (C.c_int, [C.c_uint32, C.POINTER(C.c_size_t)]),
    "vkgsb_sort_key_value_indirect": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "vkgsb_group_export": (C.c_int, [_P, _P]),
    "vkgsb_group_join": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    "vkgsb_group_join_local": (C.c_int, [C.POINTER(_P), C.c_uint32, _P]),
    "vkgsb_group_leave": (C.c_int, [_P]),
    "vkgsb_external_import": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    "vkgsb_external_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_P), C.POINTER(C.c_int), C.POINTER(_P)]),
    "vkgsb_external_release": (C.c_int, [_P]),
    "vkgsb_external_semaphore_import": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "vkgsb_external_semaphore_signal": (C.c_int, [_P, _P]),
    "vkgsb_external_semaphore_wait": (C.c_int, [_P, _P]),
    "vkgsb_external_semaphore_release": (C.c_int, [_P]),
    "vkgsb_shared_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_P), _P]),
    "vkgsb_shared_open": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "vkgsb_shared_close": (C.c_int, [C.c_int, _P]),
    "vkgsb_shared_read": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_size_t, _P]),
    "vkgsb_shared_destroy": (C.c_int, [C.c_int, _P]),
    "vkgsb_camera_orbit": (C.c_int, [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, _P,
                                     C.POINTER(CameraBlock)]),
    "vkgsb_camera_apply": (C.c_int, [C.c_uint32, C.c_uint32] + [C.c_float] * 8 + [C.POINTER(CameraBlock)]),
}

_lib = None


class VkgsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vkgsb error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m vkgs_b200.build` "
                               "(this package has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != OK:
        raise VkgsbError(code, lib().vkgsb_last_error().decode("utf-8", "replace"))
