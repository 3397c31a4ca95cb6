"""Opaque line layer of the oracle (the reference's axis / grid and the depth interaction it
exists for, engine.cc:1440-1469, 298-299).  CPU only: the rasterisation rule stated in oracle/vkgs_oracle.c
(vko_raster_lines) and the depth-tested compositing of vko_raster_rows_layer.  Parity of the rule itself is unpinned:
Vulkan leaves non-strict line rasterisation to the implementation and the reference has no test for it."""
import numpy as np

import vkgs_b200
from oracle import oracle as O
from vkgs_b200 import camera as pycam

IDENT = np.eye(4, dtype=np.float32).reshape(16)


def test_horizontal_and_vertical_lines_cover_one_pixel_per_step():
    # ndc y = 0 -> screen y = H/2 = 4.0 -> row 4; centres 0.5 .. 15.5 all inside [0, 16)
    d, c = O.raster_lines([[-1, 0, 0.5, 1, 0, 0.5]], [[1, 0, 0, 1, 1, 0, 0, 1]], IDENT, 16, 8)
    assert np.array_equal(np.argwhere(d < 1), [[4, x] for x in range(16)])
    assert np.all(d[4] == 0.5) and np.all(c[4] == [255, 0, 0, 255]) and np.all(c[3] == [0, 0, 0, 255])
    d, c = O.raster_lines([[0, -1, 0.25, 0, 1, 0.75]], [[0, 1, 0, 1, 0, 1, 0, 1]], IDENT, 16, 8)
    assert np.array_equal(np.argwhere(d < 1), [[y, 8] for y in range(8)])
    assert np.all(np.diff(d[:, 8]) > 0)                      # depth interpolates along the line


def test_nearest_line_wins_and_clipping_keeps_the_visible_part():
    lines = [[-1, 0, 0.8, 1, 0, 0.8], [-1, 0, 0.2, 1, 0, 0.2]]
    cols = [[1, 0, 0, 1, 1, 0, 0, 1], [0, 0, 1, 1, 0, 0, 1, 1]]
    d, c = O.raster_lines(lines, cols, IDENT, 16, 8)
    assert np.all(d[4] == np.float32(0.2)) and np.all(c[4] == [0, 0, 255, 255])   # depth test LESS + write
    d2, c2 = O.raster_lines(lines[::-1], cols[::-1], IDENT, 16, 8)                  # order-independent
    assert np.array_equal(d, d2) and np.array_equal(c, c2)
    # a line leaving the frustum on the right and one crossing the near plane (w changes sign) are clipped, not dropped
    d, _ = O.raster_lines([[0, 0, 0.5, 3, 0, 0.5]], [[1, 1, 1, 1, 1, 1, 1, 1]], IDENT, 16, 8)
    assert np.array_equal(np.argwhere(d < 1), [[4, x] for x in range(8, 16)])
    cam = pycam.orbit(64, 48)
    pvm = O.compose_pvm(cam.projection_matrix(), cam.view_matrix())
    eye = np.asarray(cam.eye(), np.float32)
    behind = 2 * eye + np.float32([0.5, 0.0, 0.0])            # a point behind the camera, off the line of sight
    d, _ = O.raster_lines([[0, 0, 0, *behind]], [[1, 1, 1, 1, 1, 1, 1, 1]], pvm, 64, 48)
    assert (d < 1).sum() > 0 and np.all(d[d < 1] >= 0)


def test_reference_overlay_geometry():
    pos, col, model = vkgs_b200.reference_overlay()
    assert pos.shape == (3 + 42, 2, 3) and col.shape == (45, 2, 4)        # engine.cc:618-680: 3 axes + 2 x 21 grid lines
    assert np.array_equal(model.reshape(4, 4), np.diag([10, 10, 10, 1]).astype(np.float32))
    assert np.array_equal(col[0], [[1, 0, 0, 1]] * 2) and np.all(col[3:, :, :3] == 0.5)
    p, c, _ = vkgs_b200.reference_overlay(show_grid=False)
    assert p.shape[0] == 3


def test_splats_are_depth_tested_against_the_layer_and_composited_over_it():
    w, h = 32, 16
    # one big opaque splat at depth 0.5 covering the frame: instance record {ndc xyz, pad, RS (2x2), rgb, opacity}
    inst = np.array([[0, 0, 0.5, 0, 2.0, 0, 0, 2.0, 0, 1, 0, 1.0]], np.float32)
    depth = np.ones((h, w), np.float32)
    rgba = np.zeros((h, w, 4), np.uint8); rgba[..., 3] = 255
    depth[:, :8] = 0.25; rgba[:, :8] = [255, 0, 0, 255]        # a red layer in FRONT of the splat: hides it
    depth[:, 8:16] = 0.75; rgba[:, 8:16] = [0, 0, 255, 255]    # a blue layer BEHIND it: the splat blends over blue
    for mode in (0, 1):
        plain = O.raster(inst, w, h, mode=mode)
        img = O.raster_layer(inst, w, h, depth, rgba, mode=mode)
        assert np.array_equal(img[:, :8], rgba[:, :8])
        assert np.array_equal(img[:, 16:], plain[:, 16:])      # no layer: unchanged
        mid = img[h // 2, 12].astype(int)                      # centre row: alpha ~ 1 -> the splat's green
        assert mid[1] > 200 and mid[2] < 60
        edge = img[0, 12].astype(int)                          # towards the rim the blue layer shows through
        assert edge[2] > mid[2]
