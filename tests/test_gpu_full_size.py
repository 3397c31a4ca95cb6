"""BASELINE.json full-size configuration (C2: 6,131,954 splats, 1600x900) through size-independent properties, plus
the stages the oracle still finishes in seconds at this size (cull, sort, projection)."""
import numpy as np
import pytest

import vkgs_b200
from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    rows = synth.scene_bicycle()
    r = vkgs_b200.Renderer(max_splats=rows.shape[0], max_width=1600, max_height=900)
    r.upload_splats(rows)
    r.set_option(vkgs_b200.OPT_KEEP_INSTANCES, 1)
    del rows
    yield r
    r.close()


def test_c2_visible_set_sorted_order_and_records_bit_exact(c2):
    r = c2
    cam = pycam.orbit(1600, 900, r=4.0, phi_deg=70.0, theta_deg=30.0)
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    r.set_viewport(1600, 900)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    r.set_camera(P, V, E)
    img = r.draw().copy()
    st = r.stats()
    keys, ids = r.read_sorted()
    inst = r.read_instances()
    assert st["pair_overflow"] == 0

    sc = O.Scene(*r.read_scene())
    ok, oi = O.cull(sc, O.compose_pvm(P, V))
    assert st["visible_point_count"] == len(oi) > 500_000                 # visible count: exact
    ok, oi = O.sort_pairs(ok, oi)
    assert np.array_equal(keys, ok) and np.array_equal(ids, oi)           # bit-exact order incl. ties by id
    oinst = O.project(sc, oi, O.make_camera(P, V, E, 1600, 900), 0)
    a = inst.view(np.uint32).copy(); b = oinst.view(np.uint32).copy()
    a[np.isnan(inst)] = 0; b[np.isnan(oinst)] = 0
    assert np.array_equal(a, b)                                           # records: bit-exact

    # properties
    assert np.all(keys[1:] >= keys[:-1])                                  # sortedness
    assert len(np.unique(ids)) == len(ids)                                # each visible splat exactly once
    assert np.array_equal(r.draw(), img)                                  # idempotence
    assert img[..., 3].min() >= 0 and img.shape == (900, 1600, 4)

    # band-sharded rendering (the 8-GPU partition of SURVEY §8e) reassembles the same image
    out = np.zeros_like(img)
    edges = np.linspace(0, 900, 9).astype(int)
    for y0, y1 in zip(edges[:-1], edges[1:]):
        r.set_band(int(y0), int(y1))
        out[y0:y1] = r.draw()[y0:y1]
    r.set_band(0, 0)
    assert np.array_equal(out, img)


def test_c2_image_against_oracle_on_a_crop(c2):
    """The oracle's rasteriser is O(fragments) on the CPU; a 1600x900 frame of 2 M splats takes minutes, so the
    full-size image is checked on the whole frame at reduced splat count by raising the camera far enough that the
    oracle finishes in seconds, and on mode-to-mode consistency."""
    r = c2
    cam = pycam.orbit(1600, 900, r=60.0, phi_deg=80.0, theta_deg=10.0)     # everything small and far: few fragments
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    r.set_viewport(1600, 900)
    r.set_camera(P, V, E)
    sc = O.Scene(*r.read_scene())
    for mode in (vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8):
        r.set_blend_mode(mode)
        img = r.draw().copy()
        ref = O.render(sc, O.make_camera(P, V, E, 1600, 900), mode=mode)
        d = np.abs(img.astype(np.int32) - ref["image"].astype(np.int32))
        assert d.max() <= 1, f"mode {mode}: {d.max()}/255 off"           # <= 1/255 per channel
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
