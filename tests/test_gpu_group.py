"""Band group (vkgsb_group_*, SURVEY.md 8e / BASELINE configs[4]): W renderers draw the W screen bands of the same frame,
the cull shared out between them - member j tests its 1/W of the splats against every band and writes each band's
visibility bits into that band's member (on a multi-GPU box: over NVLink), frame-numbered flags instead of a collective.
Here the members live in one process on cuda:0 (vkgsb_group_join_local); the multi-process form is what
bench.py --config c5 runs.  The assembled frame must be the ungrouped full frame, bit for bit."""
import numpy as np
import pytest

import vkgs_b200
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

pytestmark = pytest.mark.gpu

W, H, N = 800, 600, 60_000


def _member(rows):
    r = vkgs_b200.Renderer(max_splats=1 << 16, max_width=W, max_height=H, max_pairs=1 << 24)
    r.upload_splats(rows)
    r.set_viewport(W, H)
    return r


@pytest.mark.parametrize("mode", [vkgs_b200.BLEND_FP32, vkgs_b200.BLEND_UNORM8])
def test_group_of_three_assembles_the_full_frame(mode):
    rows = synth.scene_c1(n=N, seed=31)
    edges = [0, 170, 390, H]
    members = [_member(rows) for _ in range(3)]
    ref = _member(rows)
    base, _ = vkgs_b200.shared_create(0, W * H * 4)
    try:
        for r in members + [ref]:
            r.set_blend_mode(mode)
        vkgs_b200.group_join_local(members, edges)
        for i in range(5):                                   # more frames than parities: the hand-shake flags cycle
            cam = pycam.orbit(W, H, r=2.2 + 0.3 * i, phi_deg=50.0 + 7 * i, theta_deg=40.0 * i)
            P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
            ref.set_camera(P, V, E)
            full = ref.draw().copy()
            for r in members:                                # every member issues the frame, asynchronously ...
                r.set_camera(P, V, E)
                r.draw_device(dst_ptr=base)                  # ... its band's rows of ONE frame in shared memory
            for r in members:
                r.sync()
            got = vkgs_b200.shared_read(0, base, 0, (H, W, 4))
            assert np.array_equal(got, full), f"frame {i}"
            # every member culled, sorted and projected only what can reach its band
            vis = [r.stats()["visible_point_count"] for r in members]
            assert all(0 < v < ref.stats()["visible_point_count"] for v in vis)
            for g, r in enumerate(members):                  # and exactly the set the ungrouped band cull keeps
                ref.set_band(edges[g], edges[g + 1])
                ref.draw()
                assert ref.stats()["visible_point_count"] == vis[g]
            ref.set_band(0, 0)
        # leaving the group gives ordinary renderers back
        for r in members:
            r.group_leave()
        members[0].set_camera(P, V, E)
        assert np.array_equal(members[0].draw(), full)
    finally:
        vkgs_b200.shared_destroy(0, base)
        for r in members + [ref]:
            r.close()


def test_a_missing_member_is_reported_not_waited_for_forever():
    rows = synth.scene_c1(n=20_000, seed=32)
    members = [_member(rows) for _ in range(2)]
    try:
        vkgs_b200.group_join_local(members, [0, 300, H])
        cam = pycam.orbit(W, H)
        members[0].set_camera(cam.projection_matrix(), cam.view_matrix(), cam.eye())
        members[0].draw_device()                             # member 1 never issues the frame
        with pytest.raises(vkgs_b200.VkgsbError) as e:
            members[0].sync()
        assert "band group" in str(e.value)
    finally:
        for r in members:
            r.close()
