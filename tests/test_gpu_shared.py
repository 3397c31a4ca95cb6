"""Device destinations: vkgsb_draw renders straight into the caller's device memory (no copy), including memory another
process of the node owns (vkgsb_shared_*, CUDA IPC) - the delivery path of the multi-GPU view / band sharding
(bench.py --gpus N).  Two processes share cuda:0 here; on a multi-GPU box the mapped pointer lives on
another GPU and the same stores travel over NVLink."""
import multiprocessing as mp

import numpy as np
import pytest

import vkgs_b200
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu

W, H, N = 320, 192, 20_000


def _renderer():
    r = vkgs_b200.Renderer(max_splats=1 << 15, max_width=W, max_height=H, max_pairs=1 << 22)
    r.upload_splats(synth.scene_c1(n=N, seed=5))
    r.set_viewport(W, H)
    cam = pycam.orbit(W, H)
    r.set_camera(cam.projection_matrix(), cam.view_matrix(), cam.eye())
    return r


def _producer(handle, slot, band, q):
    try:
        ptr = vkgs_b200.shared_open(0, handle)
        with _renderer() as r:
            if band:
                r.set_band(*band)
            r.draw_device(dst_ptr=ptr + slot * W * H * 4)
            r.sync()
        vkgs_b200.shared_close(0, ptr)
        q.put("ok")
    except Exception as e:  # noqa: BLE001
        q.put(repr(e))


def test_device_destination_is_rendered_in_place_and_shared_across_processes():
    img_bytes = W * H * 4
    with _renderer() as r:
        ref = r.draw().copy()
        base, handle = vkgs_b200.shared_create(0, 3 * img_bytes)
        try:
            # same process: a device destination receives exactly the host image
            r.draw_device(dst_ptr=base)
            r.sync()
            assert np.array_equal(vkgs_b200.shared_read(0, base, 0, (H, W, 4)), ref)
            # another process renders a whole view into slot 1 and two bands of one frame into slot 2
            ctx = mp.get_context("spawn")
            q = ctx.Queue()
            jobs = [(1, None), (2, (0, 80)), (2, (80, H))]
            for slot, band in jobs:
                p = ctx.Process(target=_producer, args=(handle, slot, band, q))
                p.start()
                assert q.get(timeout=300) == "ok"
                p.join(timeout=60)
            assert np.array_equal(vkgs_b200.shared_read(0, base, img_bytes, (H, W, 4)), ref)
            assert np.array_equal(vkgs_b200.shared_read(0, base, 2 * img_bytes, (H, W, 4)), ref)   # bands assemble in place
        finally:
            vkgs_b200.shared_destroy(0, base)


def test_batch_to_device_destination():
    with _renderer() as r:
        cams = []
        for i in range(3):
            c = pycam.orbit(W, H, theta_deg=40.0 * i)
            cams.append(vkgs_b200.camera_block(c.projection_matrix(), c.view_matrix(), c.eye()))
        host = r.draw_batch(cams)
        base, _ = vkgs_b200.shared_create(0, 3 * W * H * 4)
        try:
            r.draw_batch(cams, dst_ptr=base)
            r.sync()
            assert np.array_equal(vkgs_b200.shared_read(0, base, 0, (3, H, W, 4)), host)
        finally:
            vkgs_b200.shared_destroy(0, base)
