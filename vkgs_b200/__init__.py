"""vkgs_b200: B200-native Gaussian-splat renderer path behind the jaesung-cs/vkgs renderer surface.

The compute lives in vkgs_b200/lib/libvkgsb.so (hand-written sm_100a kernels behind the C ABI of include/vkgsb.h);
this package is the thin host side.  There is no CPU / PyTorch fallback: using the renderer without the built
library, or without a CUDA device, raises.
"""
from . import camera, synth  # noqa: F401
from ._lib import (BLEND_FP32, BLEND_UNORM8, EXTERNAL_CUDA_POSIX_FD, EXTERNAL_OPAQUE_FD, FORMAT_BGRA8, FORMAT_RGBA8, LIB_PATH, OPT_BAND_CULL, OPT_BAND_Y0,  # noqa: F401
                   OPT_COUNT_FRAGMENTS, OPT_UNORM8_CUT_EXP, OPT_L2_PIN_MB, OPT_SPATIAL_ORDER,
                   OPT_BAND_Y1, OPT_KEEP_INSTANCES, OPT_BLEND_MODE, OPT_PIXEL_FORMAT, OPT_STAGE_TIMING, CameraBlock, VkgsbError)
from .renderer import (Renderer, camera_block, device_count, external_alloc, external_import, external_release,
                       group_join_local, orbit_camera_block, reference_overlay,  # noqa: F401
                       shared_close, shared_create, shared_destroy, shared_open, shared_read, sort_key_value_indirect,
                       sort_storage_bytes)

__all__ = ["Renderer", "camera_block", "orbit_camera_block", "device_count", "sort_storage_bytes",
           "sort_key_value_indirect", "reference_overlay", "camera", "synth", "VkgsbError", "CameraBlock"]
