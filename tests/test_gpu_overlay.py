"""Opaque line layer on the GPU (vkgsb_set_lines: the reference's axis / grid drawn under the splats with depth test +
write, the splats depth-tested LESS against it; engine.cc:1440-1469, 298-299) against the oracle's statement of the
same rule: the layer bit-exact through its effect on the image, the composited frame within 1/255."""
import numpy as np
import pytest

import vkgs_b200
from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    rows = synth.scene_c1(60_000, seed=808)
    r = vkgs_b200.Renderer(max_splats=1 << 16, max_width=800, max_height=600, max_pairs=1 << 24)
    r.set_option(vkgs_b200.OPT_KEEP_INSTANCES, 1)
    r.upload_splats(rows)
    yield r
    r.close()


def _oracle_with_layer(r, P, V, E, w, h, pos, col, model, mode):
    sc = O.Scene(*r.read_scene())
    cam = O.make_camera(P, V, E, w, h)
    ref = O.render(sc, cam, mode=mode)
    pvm_lines = O.compose_pvm(P, V, np.asarray(model, np.float32).reshape(4, 4).T)
    depth, rgba = O.raster_lines(pos, col, pvm_lines, w, h)
    return O.raster_layer(ref["inst"], w, h, depth, rgba, mode=mode), ref["image"], depth, rgba


@pytest.mark.parametrize("mode", [vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8])
@pytest.mark.parametrize("view", [(2.0, 45.0, 45.0), (0.7, 80.0, 200.0), (6.0, 20.0, 10.0)])
def test_reference_axis_and_grid_under_the_splats(scene, mode, view):
    r = scene
    w, h = 800, 600
    cam = pycam.orbit(w, h, r=view[0], phi_deg=view[1], theta_deg=view[2])   # inside the grid, near an axis, far away
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    pos, col, model = vkgs_b200.reference_overlay()
    r.set_viewport(w, h)
    r.set_blend_mode(mode)
    r.set_camera(P, V, E)
    r.set_lines(None)
    plain = r.draw().copy()
    r.set_lines(pos, col, model)
    img = r.draw().copy()
    want, want_plain, depth, rgba = _oracle_with_layer(r, P, V, E, w, h, pos, col, model, mode)
    assert (depth < 1).sum() > 500                                            # the overlay is on screen
    assert np.abs(plain.astype(int) - want_plain.astype(int)).max() <= 1      # without the layer: as before
    d = np.abs(img.astype(int) - want.astype(int))
    assert d.max() <= 1, f"{(d > 1).sum()} pixels off by up to {d.max()}/255"  # <= 1/255 per channel
    off = depth >= 1
    assert np.array_equal(img[off], plain[off])                               # pixels no line touches are unchanged
    assert not np.array_equal(img, plain)
    assert np.array_equal(r.draw(), img)                                      # idempotent (graph replay)
    r.set_lines(None)
    assert np.array_equal(r.draw(), plain)                                    # and it can be removed again


def test_lines_alone_match_the_oracle_layer_exactly(scene):
    """With every splat culled (camera looking away from the scene) the frame is the layer itself: the GPU line
    rasteriser must pick the same pixels, depths' winners and colours as the oracle's rule, bit for bit."""
    r = scene
    w, h = 640, 480
    r.set_viewport(w, h)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    rng = np.random.default_rng(5)
    n = 300
    pos = rng.uniform(-1.5, 1.5, (n, 2, 3)).astype(np.float32)
    pos[:, :, 2] += 40.0                                       # lines around z = 40, the splats are around the origin
    col = rng.uniform(0, 1, (n, 2, 4)).astype(np.float32)
    col[: n // 2, :, 3] = 1.0
    cam = pycam.orbit(w, h, r=4.0, phi_deg=90.0, theta_deg=0.0, center=(0.0, 0.0, 40.0))   # eye at z = 44, looking at -z
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    r.set_camera(P, V, E)
    r.set_lines(pos, col, None)
    img = r.draw().copy()
    depth, rgba = O.raster_lines(pos, col, O.compose_pvm(P, V), w, h)
    vis = r.stats()["visible_point_count"]
    if vis == 0:
        assert np.array_equal(img, rgba)
    else:                                                      # some splats in view: compare where none can reach
        want, _, _, _ = _oracle_with_layer(r, P, V, E, w, h, pos, col, np.eye(4, dtype=np.float32).reshape(16), 0)
        assert np.abs(img.astype(int) - want.astype(int)).max() <= 1
    assert (depth < 1).sum() > 1000
    r.set_lines(None)
