"""External-memory destinations (vkgsb_external_*): memory another API / process owns, handed over
as a file descriptor, becomes a frame destination - the reference's interop pattern (interop/cuda_image.cu:77-132) with
the VkImage's memory replaced by the one fd exporter that exists without a Vulkan loader: a CUDA virtual-memory
allocation (cuMemExportToShareableHandle).  The frame written through the IMPORTED mapping must be readable through the
OWNER's mapping."""
import os

import numpy as np
import pytest

import vkgs_b200
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu

W, H = 320, 192


def test_frame_lands_in_memory_imported_by_file_descriptor():
    nbytes = W * H * 4
    owner, fd, owner_ptr = vkgs_b200.external_alloc(0, nbytes)
    try:
        imported, ptr = vkgs_b200.external_import(0, fd, nbytes, vkgs_b200.EXTERNAL_CUDA_POSIX_FD)
        try:
            assert ptr != owner_ptr                                   # a second mapping of the same memory
            with vkgs_b200.Renderer(max_splats=1 << 15, max_width=W, max_height=H, max_pairs=1 << 22) as r:
                r.upload_splats(synth.scene_c1(n=20_000, seed=41))
                r.set_viewport(W, H)
                cam = pycam.orbit(W, H)
                r.set_camera(cam.projection_matrix(), cam.view_matrix(), cam.eye())
                ref = r.draw().copy()
                r.draw_device(dst_ptr=ptr)                            # the blend kernel writes through the imported mapping
                r.sync()
                assert np.array_equal(vkgs_b200.shared_read(0, owner_ptr, 0, (H, W, 4)), ref)
        finally:
            vkgs_b200.external_release(imported)
    finally:
        os.close(fd)
        vkgs_b200.external_release(owner)


def test_bad_descriptors_are_errors_not_crashes():
    with pytest.raises(vkgs_b200.VkgsbError):
        vkgs_b200.external_import(0, -1, 4096)
    r, w = os.pipe()                                                  # a descriptor that is no memory object
    try:
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            vkgs_b200.external_import(0, r, 4096, vkgs_b200.EXTERNAL_CUDA_POSIX_FD)
        assert e.value.code == 2                                      # VKGSB_ERR_CUDA
    finally:
        os.close(r); os.close(w)
