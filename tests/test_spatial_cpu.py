"""The tile-box decisions of the cull (k_cull_classify / classify_tile in vkgs_b200/csrc/project.cu, restated in numpy
float32 by tools/spatial_model.py) against the per-splat test they stand for: a tile classified "outside" must hold no
visible splat of the oracle's cull (rank.comp:31-41), a tile classified "inside" only visible ones, and a tile the band
test drops no splat the per-splat band bound keeps.  The stored order itself (Morton code of the centres) must be a
permutation that makes tiles of 256 splats compact."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import spatial_model as SM  # noqa: E402

VIEWS = [(1.5, 70.0, 30.0, 60.0), (0.4, 85.0, 200.0, 60.0), (8.0, 20.0, 10.0, 35.0), (45.0, 60.0, 300.0, 100.0)]


@pytest.fixture(scope="module")
def ordered_scene():
    rows = synth.scene_bicycle(300_000, seed=77)
    sc = O.activate(rows, synth.STANDARD_OFFSETS)
    lmax = np.exp(2.0 * rows[:, synth._COL["scale_0"]:synth._COL["scale_2"] + 1].max(1))   # largest eigenvalue of Sigma
    order = SM.spatial_order(sc.pos, lmax)                        # large splats first, Morton order inside each class
    assert np.array_equal(np.sort(order), np.arange(len(rows)))
    assert lmax[order[:256]].min() > np.median(lmax)
    return O.Scene(sc.pos[order], sc.cov[order], sc.opacity[order], sc.sh[order])


def test_stored_order_makes_tiles_compact(ordered_scene):
    pos = ordered_scene.pos
    lo, hi = SM.tile_boxes(pos)
    ext = (hi - lo).max(1)
    assert np.median(ext) < 0.1 * (pos.max(0) - pos.min(0)).max()


@pytest.mark.parametrize("view", VIEWS)
def test_box_decisions_imply_the_per_splat_cull(ordered_scene, view):
    sc = ordered_scene
    cam = pycam.orbit(1600, 900, r=view[0], phi_deg=view[1], theta_deg=view[2], fovy_deg=view[3])
    pvm = O.compose_pvm(cam.projection_matrix(), cam.view_matrix())
    _, ids = O.cull(sc, pvm)
    vis = np.zeros(len(sc.pos), bool)
    vis[ids] = True
    lo, hi = SM.tile_boxes(sc.pos)
    cls = SM.classify(pvm, lo, hi)
    nt = len(lo)
    pad = nt * 256 - len(vis)
    vt = np.concatenate([vis, np.zeros(pad, bool)]).reshape(nt, 256).sum(1)
    full = np.full(nt, 256)
    full[-1] -= pad
    assert not ((cls == 0) & (vt != 0)).any(), "a tile decided 'outside' holds a visible splat"
    assert not ((cls == 1) & (vt != full)).any(), "a tile decided 'inside' holds an invisible splat"
    if 0 < len(ids) < len(vis):
        assert (cls != 2).mean() > 0.5, "the boxes decide most tiles"


@pytest.mark.parametrize("view", VIEWS[:3])
def test_band_decision_drops_no_splat_the_per_splat_bound_keeps(ordered_scene, view):
    sc = ordered_scene
    w, h = 1600, 900
    cam = pycam.orbit(w, h, r=view[0], phi_deg=view[1], theta_deg=view[2], fovy_deg=view[3])
    P, V = cam.projection_matrix(), cam.view_matrix()
    pvm = O.compose_pvm(P, V)
    cov = sc.cov.astype(np.float64)
    S = np.empty((len(cov), 3, 3))
    S[:, 0, 0], S[:, 1, 0], S[:, 2, 0], S[:, 1, 1], S[:, 2, 1], S[:, 2, 2] = cov.T
    S[:, 0, 1], S[:, 0, 2], S[:, 1, 2] = S[:, 1, 0], S[:, 2, 0], S[:, 2, 1]
    lmax = np.linalg.eigvalsh(S)[:, -1].astype(np.float32) * np.float32(1.0001)
    lo, hi = SM.tile_boxes(sc.pos)
    nt = len(lo)
    pad = nt * 256 - len(lmax)
    trmax = np.concatenate([lmax, np.zeros(pad, np.float32)]).reshape(nt, 256).max(1)
    bc = SM.band_params(P, V, w, h)
    _, ids = O.cull(sc, pvm)
    vis = np.zeros(len(sc.pos), bool)
    vis[ids] = True
    dropped_any = False
    for y0, y1 in [(0, 113), (113, 450), (450, 451), (787, 900)]:
        miss = SM.band_miss_splats(pvm, sc.pos, lmax, bc, h, y0, y1)
        keep = vis & ~miss                                            # what the per-splat cull keeps for the band
        kt = np.concatenate([keep, np.zeros(pad, bool)]).reshape(nt, 256).any(1)
        tile_miss = SM.band_classify(pvm, lo, hi, trmax, bc, h, y0, y1)
        assert not (tile_miss & kt).any(), f"band [{y0},{y1}): a dropped tile holds a splat the per-splat test keeps"
        dropped_any |= bool(tile_miss.any())
    assert dropped_any
