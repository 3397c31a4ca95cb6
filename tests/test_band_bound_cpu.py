"""The bound behind VKGSB_OPT_BAND_CULL (renderer.cu fill_params / project.cu band_miss), checked on the CPU against the
oracle's projection: for every visible splat the footprint's pixel half-height
    ey = 3 * (H/2) * (|RS10| + |RS11|)                                      (raster_record / splat.vert:10-26)
must not exceed sqrt(9 (H/2)^2 ((P00^2 + P11^2 + x^2 + y^2) |W|_2^2 lambda_max(Sigma) / w^2 + lpx + lpy)),
or a band could lose a splat that reaches it."""
import numpy as np
import pytest

from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import synth


@pytest.mark.parametrize("view", [(2.0, 45.0, 45.0, 60.0), (0.6, 85.0, 200.0, 60.0), (8.0, 20.0, 10.0, 35.0), (1.2, 60.0, 300.0, 100.0)])
@pytest.mark.parametrize("model_scale", [1.0, 2.5])
def test_footprint_half_height_never_exceeds_the_band_cull_bound(view, model_scale):
    w, h = 1280, 720
    rows = synth.scene_c1(40_000, seed=99)
    rows[:2000, synth._COL["scale_0"]:synth._COL["scale_2"] + 1] += 2.5     # some large and strongly anisotropic splats
    rows[2000:4000, synth._COL["scale_0"]] += 3.0
    sc = O.activate(rows, synth.STANDARD_OFFSETS)
    cam = pycam.orbit(w, h, r=view[0], phi_deg=view[1], theta_deg=view[2], fovy_deg=view[3])
    P, V, E = cam.projection_matrix(), cam.view_matrix(), cam.eye()
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = model_scale * np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1]], np.float32)   # rotation * scale
    pvm = O.compose_pvm(P, V, M)
    keys, ids = O.cull(sc, pvm)
    assert len(ids) > 5_000
    inst = O.project(sc, ids, O.make_camera(P, V, E, w, h, M), 0)
    hh = 0.5 * h
    ey = 3.0 * hh * (np.abs(inst[:, 5].astype(np.float64)) + np.abs(inst[:, 7].astype(np.float64)))

    # the bound, as fill_params / band_miss compute it (float64 here: the kernel adds 1 % and 2 pixels of slack)
    Pm = np.asarray(P, np.float64).reshape(4, 4).T               # column-major -> [row, col]
    assert Pm[3, 2] == -1.0 and Pm[0, 1] == 0.0 and Pm[1, 0] == 0.0   # the projection shape the bound assumes
    W = (np.asarray(V, np.float64).reshape(4, 4).T[:3, :3]) @ M[:3, :3].astype(np.float64)
    w2 = np.linalg.norm(W, 2) ** 2
    cov = sc.cov[ids].astype(np.float64)                          # c00 c10 c20 c11 c21 c22
    S = np.empty((len(ids), 3, 3))
    S[:, 0, 0], S[:, 1, 0], S[:, 2, 0], S[:, 1, 1], S[:, 2, 1], S[:, 2, 2] = cov.T
    S[:, 0, 1], S[:, 0, 2], S[:, 1, 2] = S[:, 1, 0], S[:, 2, 0], S[:, 2, 1]
    lmax = np.linalg.eigvalsh(S)[:, -1]
    pos = sc.pos[ids].astype(np.float64)
    clip = (np.asarray(pvm, np.float64).reshape(4, 4).T @ np.c_[pos, np.ones(len(ids))].T).T
    x, y, iw = clip[:, 0] / clip[:, 3], clip[:, 1] / clip[:, 3], 1.0 / clip[:, 3]
    bound2 = 9.0 * hh * hh * ((Pm[0, 0] ** 2 + Pm[1, 1] ** 2 + x * x + y * y) * w2 * lmax * iw * iw + 1.0 / w ** 2 + 1.0 / h ** 2)
    ok = np.isfinite(ey)
    assert ok.sum() > 0.99 * len(ey)
    slack = np.sqrt(bound2[ok]) * 1.005 + 0.5 - ey[ok]
    assert slack.min() >= 0.0, f"bound violated by {-slack.min():.3f} px"
    # and it is a useful bound, not a vacuous one: within a small factor of the true extent for most splats
    assert np.median(np.sqrt(bound2[ok]) / np.maximum(ey[ok], 1e-9)) < 3.0
