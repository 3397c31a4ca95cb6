"""GPU parity tests: the CUDA path, called through the C ABI (include/vkgsb.h via ctypes), against
  (a) the CPU oracle on the same inputs - bit-exact for visible count, keys, ids and instance records,
      <= 1/255 per channel for the image (tolerance written in each test);
  (b) the committed golden fixtures generated from the reference's own shaders (tests/golden/make_golden.py).
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import vkgs_b200
from conftest import GOLDEN_NAMES, load_golden, ulp_diff
from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def bits(a):
    """Bit pattern with every NaN canonicalised (x86 produces 0xFFC00000 for 0/0, the GPU 0x7FFFFFFF)."""
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


@pytest.fixture(scope="module")
def small_renderer():
    r = vkgs_b200.Renderer(max_splats=1 << 17, max_width=1024, max_height=768, max_pairs=1 << 24)
    r.set_option(vkgs_b200.OPT_KEEP_INSTANCES, 1)      # parity tap: reference-format instance records
    yield r
    r.close()


def device_scene(r):
    pos, cov, op, sh = r.read_scene()
    return O.Scene(pos, cov, op, sh)


def oracle_frame(scene, proj, view, eye, w, h, model, mode):
    cam = O.make_camera(proj, view, eye, w, h, model)
    return O.render(scene, cam, mode=mode), cam


# ------------------------------------------------------------------------------------------------ activation (a2)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_activation_against_golden_and_oracle(small_renderer, name):
    g = load_golden(name)
    r = small_renderer
    r.upload_splats(g["rows"], g["offsets"])
    pos, cov, op, sh = r.read_scene()
    order = r.read_order()                                   # stored splat i = splat order[i] of the file (spatial order)
    assert np.array_equal(np.sort(order), np.arange(len(g["pos"])))
    assert np.array_equal(pos, g["pos"][order])
    assert np.array_equal(sh, g["sh"][order])                # f16 RNE bits
    assert ulp_diff(op, g["opacity"][order]).max() <= 2      # device expf vs libm
    scale = np.abs(g["cov"][order]).max(axis=1, keepdims=True)
    assert (np.abs(cov - g["cov"][order]) / scale).max() < 4e-6
    o = O.activate(g["rows"], g["offsets"])
    assert (np.abs(cov - o.cov[order]) / scale).max() < 1e-6  # same operation order; only exp differs


# ------------------------------------------------------------------------- whole frame on the golden fixtures
@pytest.mark.parametrize("spatial", [1, 0])
@pytest.mark.parametrize("mode", [vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_frame_bit_exact_vs_oracle_and_close_to_golden(small_renderer, name, mode, spatial):
    g = load_golden(name)
    r = small_renderer
    w, h = int(g["width"]), int(g["height"])
    r.set_option(vkgs_b200.OPT_SPATIAL_ORDER, spatial)
    try:
        r.upload_splats(g["rows"], g["offsets"])
    finally:
        r.set_option(vkgs_b200.OPT_SPATIAL_ORDER, 1)
    r.set_viewport(w, h)
    r.set_blend_mode(mode)
    r.set_camera(g["proj"], g["view"], g["eye"], g["model"])
    img = r.draw()
    st = r.stats()
    keys, ids = r.read_sorted()
    inst = r.read_instances()

    ref, _ = oracle_frame(device_scene(r), g["proj"], g["view"], g["eye"], w, h, g["model"], mode)
    assert st["visible_point_count"] == ref["stats"]["visible"] == len(g["sorted_index"])   # visible count: exact
    assert np.array_equal(keys, ref["keys"])                                                # sorted keys: bit-exact
    assert np.array_equal(ids, ref["ids"])                                                  # sorted ids: bit-exact
    assert np.array_equal(bits(inst), bits(ref["inst"]))                                    # 12-float records: bit-exact
    d = np.abs(img.astype(np.int32) - ref["image"].astype(np.int32))
    assert d.max() <= 1, f"image differs from the oracle by {d.max()}/255"                  # <= 1/255 per channel
    assert psnr(img, ref["image"]) > 50.0

    # against the reference shaders' own output (fixture): same visible set, records to a few ulp, image <= 1/255
    assert np.array_equal(np.sort(r.read_order()[ids]), np.sort(g["sorted_index"]))       # ids index the stored order
    # the fixture's image resolved equal depth keys (5 / 7 / 26 pairs in the three scenes; a race in the reference,
    # rank.comp:38) in file order: compared in file order
    if mode == vkgs_b200.BLEND_FP32 and not spatial:
        q = np.clip(np.rint(g["image_f32"] * 255.0), 0, 255).astype(np.int32)
        dg = np.abs(img.astype(np.int32) - q)
        assert dg.max() <= 1 and psnr(img, q) > 50.0


# ------------------------------------------------------------------------------- C1: 100k splats, 800x600
@pytest.fixture(scope="module")
def c1():
    rows = synth.scene_c1()
    cam = pycam.orbit(800, 600)
    return rows, cam.projection_matrix(), cam.view_matrix(), cam.eye()


@pytest.mark.parametrize("mode", [vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8])
def test_c1_config(small_renderer, c1, mode):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(800, 600)
    r.set_blend_mode(mode)
    r.set_camera(P, V, E)
    img = r.draw()
    keys, ids = r.read_sorted()
    inst = r.read_instances()
    ref, _ = oracle_frame(device_scene(r), P, V, E, 800, 600, None, mode)
    assert len(ids) == ref["stats"]["visible"] > 50_000
    assert np.array_equal(keys, ref["keys"]) and np.array_equal(ids, ref["ids"])
    assert np.array_equal(bits(inst), bits(ref["inst"]))
    d = np.abs(img.astype(np.int32) - ref["image"].astype(np.int32))
    assert d.max() <= 1 and psnr(img, ref["image"]) > 50.0                                  # <= 1/255, PSNR > 50 dB
    assert r.stats()["pair_overflow"] == 0


def test_graph_and_eager_paths_agree_and_are_idempotent(small_renderer, c1):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(800, 600)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    r.set_camera(P, V, E)
    r.set_option(vkgs_b200.OPT_STAGE_TIMING, 0)
    a = r.draw().copy()
    b = r.draw().copy()
    r.set_option(vkgs_b200.OPT_STAGE_TIMING, 1)
    c = r.draw().copy()
    st = r.stats()
    r.set_option(vkgs_b200.OPT_STAGE_TIMING, 0)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert st["ms_total"] > 0 and st["ms_project"] > 0 and st["ms_sort"] > 0 and st["ms_blend"] > 0


def test_band_rendering_concatenates_to_the_full_image(small_renderer, c1):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(800, 600)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    r.set_camera(P, V, E)
    r.set_band(0, 0)
    full = r.draw().copy()
    out = np.zeros_like(full)
    for y0, y1 in ((0, 150), (150, 290), (290, 600)):     # not tile aligned on purpose
        r.set_band(y0, y1)
        out[y0:y1] = r.draw()[y0:y1]
    r.set_band(0, 0)
    assert np.array_equal(out, full)


def test_bgra_is_a_channel_swap(small_renderer, c1):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(320, 200)
    r.set_camera(P, V, E)
    a = r.draw().copy()
    r.set_option(vkgs_b200.OPT_PIXEL_FORMAT, vkgs_b200.FORMAT_BGRA8)
    b = r.draw().copy()
    r.set_option(vkgs_b200.OPT_PIXEL_FORMAT, vkgs_b200.FORMAT_RGBA8)
    assert np.array_equal(a[..., [2, 1, 0, 3]], b)


# ------------------------------------------------------------------------------------------------- edge cases
def test_nothing_visible_gives_the_clear_colour(small_renderer, c1):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows[:1000])
    r.set_viewport(64, 48)
    cam = pycam.orbit(64, 48, r=50.0)
    cam.center = np.array([0, 0, 200.0], np.float32)       # look away from the scene
    r.set_camera(cam.projection_matrix(), cam.view_matrix(), cam.eye())
    img = r.draw()
    assert r.stats()["visible_point_count"] == 0
    assert np.array_equal(img, np.broadcast_to(np.array([0, 0, 0, 255], np.uint8), img.shape))   # clear (0,0,0,1)
    k, i = r.read_sorted()
    assert len(k) == 0 and len(i) == 0


@pytest.mark.parametrize("n", [1, 31, 257, 1023, 1025, 4097])
def test_ragged_sizes(small_renderer, c1, n):
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows[:n])
    r.set_viewport(200, 120)
    cam = pycam.orbit(200, 120)
    r.set_camera(cam.projection_matrix(), cam.view_matrix(), cam.eye())
    img = r.draw()
    keys, ids = r.read_sorted()
    ref, _ = oracle_frame(device_scene(r), cam.projection_matrix(), cam.view_matrix(), cam.eye(), 200, 120, None, 0)
    assert np.array_equal(keys, ref["keys"]) and np.array_equal(ids, ref["ids"])
    assert np.abs(img.astype(np.int32) - ref["image"].astype(np.int32)).max() <= 1


def test_degenerate_splats_vanish_like_nan_lanes(small_renderer):
    """D == 0 (isotropic footprint on the optical axis) makes projection.comp:128-129 divide 0/0: the reference
    draws nothing for that splat (SURVEY.md §7 hard part 6).  Also zero quaternion, huge and tiny scales."""
    rows = synth.scene_c1(n=64, seed=5)
    c = {p: i for i, p in enumerate(synth.PLY_PROPS)}
    rows[:, 0:3] = 0.0
    rows[:8, c["scale_0"]:c["scale_2"] + 1] = -3.0                      # isotropic at the look-at point: D == 0
    rows[:8, c["rot_0"]] = 1.0; rows[:8, c["rot_1"]:c["rot_3"] + 1] = 0.0
    rows[8:16, c["rot_0"]:c["rot_3"] + 1] = 0.0                         # zero quaternion: NaN covariance
    rows[16:24, c["scale_0"]:c["scale_2"] + 1] = 30.0                   # exp(30)^2 overflows
    rows[24:32, c["scale_0"]:c["scale_2"] + 1] = -60.0                  # underflows to 0
    rows[32:, 0:3] = np.random.default_rng(3).normal(0, 0.3, (32, 3))
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(96, 96)
    cam = pycam.orbit(96, 96)
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    r.set_camera(P, V, E)
    for mode in (vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8):
        r.set_blend_mode(mode)
        img = r.draw()
        inst = r.read_instances()
        ref, _ = oracle_frame(device_scene(r), P, V, E, 96, 96, None, mode)
        assert np.array_equal(bits(inst), bits(ref["inst"]))           # NaNs in the same lanes, same payloads
        assert np.isnan(inst[:, 4:8]).any()
        assert np.abs(img.astype(np.int32) - ref["image"].astype(np.int32)).max() <= 1
    r.set_blend_mode(vkgs_b200.BLEND_FP32)


def test_errors_are_reported_not_swallowed(c1):
    rows, P, V, E = c1
    with vkgs_b200.Renderer(max_splats=2048, max_width=128, max_height=128, max_pairs=1 << 16) as r:
        r.set_viewport(64, 64)
        r.set_camera(P, V, E)
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            r.draw()
        assert e.value.code == 5                                        # VKGSB_ERR_NO_SCENE
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            r.upload_splats(rows[:4096])
        assert e.value.code == 4                                        # VKGSB_ERR_CAPACITY
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            r.set_viewport(4096, 4096)
        assert e.value.code == 4
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            r.load_ply("/nonexistent/file.ply")
        assert e.value.code == 3                                        # VKGSB_ERR_IO


def test_pair_overflow_is_flagged_and_drops_the_farthest(c1):
    rows, P, V, E = c1
    with vkgs_b200.Renderer(max_splats=1 << 17, max_width=800, max_height=600, max_pairs=4096) as r:
        r.upload_splats(rows)
        r.set_viewport(800, 600)
        r.set_camera(P, V, E)
        r.draw()
        st = r.stats()
        assert st["pair_overflow"] == 1 and st["pair_count"] == 4096


# ---------------------------------------------------------------------------------------------- PLY ingest (a2)
def test_ply_file_load_equals_upload(tmp_path, small_renderer, c1):
    rows, P, V, E = c1
    sub = rows[:70_000]                                                  # > one 65 536-vertex chunk
    path = str(tmp_path / "scene.ply")
    synth.write_ply(path, sub)
    r = small_renderer
    r.upload_splats(sub)
    a = r.read_scene()
    r.load_ply(path)
    b = r.read_scene()
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    pr = r.load_progress()
    assert pr["total"] == pr["loaded"] == 70_000 and pr["state"] == 2


def test_ply_property_order_is_looked_up_by_name(tmp_path, small_renderer):
    rows = synth.scene_c1(n=3000, seed=9)
    perm = np.random.default_rng(1).permutation(len(synth.PLY_PROPS))
    props = [synth.PLY_PROPS[i] for i in perm]
    path = str(tmp_path / "shuffled.ply")
    synth.write_ply(path, rows[:, perm], props)
    r = small_renderer
    r.load_ply(path)
    a = r.read_scene()
    r.upload_splats(rows)
    b = r.read_scene()
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_async_load_progress_and_supersede(tmp_path, small_renderer, c1):
    rows, P, V, E = c1
    p1, p2 = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    synth.write_ply(p1, rows)
    synth.write_ply(p2, rows[:5000])
    r = small_renderer
    r.load_ply_async(p1)
    r.load_ply_async(p2)                                                # cancels / supersedes the first (engine.cc:541-544)
    r.wait_load()
    pr = r.load_progress()
    assert pr["state"] == 2 and pr["total"] == 5000 and pr["loaded"] == 5000
    assert r.read_scene()[0].shape[0] == 5000


def test_spatial_order_is_a_permutation_that_keeps_the_frame(small_renderer, c1):
    """VKGSB_OPT_SPATIAL_ORDER: the stored order is the file's order sorted by the Morton code of the centres; the visible
    set and the sorted keys are the same either way, the image differs only where equal depth keys swap (<= 1/255 here)."""
    rows, P, V, E = c1
    r = small_renderer
    r.set_viewport(800, 600)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    r.set_camera(P, V, E)
    r.set_option(vkgs_b200.OPT_SPATIAL_ORDER, 0)
    try:
        r.upload_splats(rows)
        assert np.array_equal(r.read_order(), np.arange(len(rows)))
        file_scene = r.read_scene()
        img0 = r.draw().copy()
        keys0, ids0 = r.read_sorted()
        ref, _ = oracle_frame(O.Scene(*file_scene), P, V, E, 800, 600, None, 0)    # all tiles are tested per splat
        assert np.array_equal(keys0, ref["keys"]) and np.array_equal(ids0, ref["ids"])
    finally:
        r.set_option(vkgs_b200.OPT_SPATIAL_ORDER, 1)
    r.upload_splats(rows)
    order = r.read_order()
    assert np.array_equal(np.sort(order), np.arange(len(rows))) and not np.array_equal(order, np.arange(len(rows)))
    stored = r.read_scene()
    for a, b in zip(stored, file_scene):
        assert np.array_equal(a.view(np.uint8), b[order].view(np.uint8))
    # neighbours in the stored order are neighbours in space: tiles of 256 are far smaller than the scene
    pos = stored[0]
    nt = len(pos) // 256
    ext = (pos[:nt * 256].reshape(nt, 256, 3).max(1) - pos[:nt * 256].reshape(nt, 256, 3).min(1)).max(1)
    assert np.median(ext) < 0.25 * (pos.max(0) - pos.min(0)).max()
    img1 = r.draw().copy()
    keys1, ids1 = r.read_sorted()
    assert np.array_equal(keys1, keys0)
    assert np.array_equal(np.sort(order[ids1]), np.sort(ids0))
    assert np.abs(img1.astype(np.int32) - img0.astype(np.int32)).max() <= 1


def test_malformed_ply_is_rejected(tmp_path, small_renderer):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    with pytest.raises(vkgs_b200.VkgsbError) as e:
        small_renderer.load_ply(str(p))
    assert e.value.code == 3
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nend_header\n\0\0\0\0")
    with pytest.raises(vkgs_b200.VkgsbError) as e:
        small_renderer.load_ply(str(p))
    assert e.value.code == 3 and "missing" in str(e.value)
    rows = synth.scene_c1(n=100, seed=2)
    good = tmp_path / "trunc.ply"
    synth.write_ply(str(good), rows)
    data = good.read_bytes()
    good.write_bytes(data[:-1000])                                       # truncated body
    with pytest.raises(vkgs_b200.VkgsbError) as e:
        small_renderer.load_ply(str(good))
    assert e.value.code == 3


def test_instances_tap_needs_the_option(c1):
    rows, P, V, E = c1
    with vkgs_b200.Renderer(max_splats=4096, max_width=128, max_height=128, max_pairs=1 << 18) as r:
        r.upload_splats(rows[:4000])
        r.set_viewport(128, 96)
        r.set_camera(P, V, E)
        r.draw()
        with pytest.raises(vkgs_b200.VkgsbError):
            r.read_instances()                                           # not kept by default
        r.set_option(vkgs_b200.OPT_KEEP_INSTANCES, 1)
        a = r.draw().copy()
        assert r.read_instances().shape[1] == 12
        r.set_option(vkgs_b200.OPT_KEEP_INSTANCES, 0)
        assert np.array_equal(r.draw(), a)                               # the tap does not change the image


def test_row_histogram_counts_the_centres_of_the_last_frame(small_renderer, c1):
    """vkgsb_row_histogram (the load screen-band partitions balance on) against the instance records of the same frame."""
    rows, P, V, E = c1
    r = small_renderer
    r.upload_splats(rows)
    r.set_viewport(800, 600)
    r.set_blend_mode(vkgs_b200.BLEND_FP32)
    r.set_camera(P, V, E)
    r.draw()
    hist = r.row_histogram()
    inst = r.read_instances()
    cpy = np.float32(300.0) * inst[:, 1] + np.float32(299.5)          # fma(ndc.y, H/2, H/2 - 1/2) up to one rounding
    want = np.bincount(np.clip(np.rint(cpy[~np.isnan(cpy)]), 0, 599).astype(np.int64), minlength=600)
    assert hist.shape == (600,) and hist.sum() == (~np.isnan(cpy)).sum() == r.stats()["visible_point_count"]
    assert np.abs(hist.astype(np.int64) - want).sum() <= 4            # a centre exactly between two rows may round either way
