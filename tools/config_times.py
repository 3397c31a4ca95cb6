#!/usr/bin/env python
"""Stage times of BASELINE.json configs[2..4] on one GPU (configs[1] is bench.py): one JSON line per config.
  C3  garden-shaped 5,834,734 splats, 1920x1080, zoomed-out camera (blend-bound)
  C4  3840x2160 orbit views of the C2 scene (one GPU's share of the 360-view batch)
  C5  50 M-splat scene, 1600x900: the whole frame, and one of 8 screen bands (what each of 8 GPUs would run)
usage: python tools/config_times.py [c3] [c4] [c5] [--n5 50000000]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkgs_b200  # noqa: E402
from vkgs_b200 import _lib as L  # noqa: E402
from vkgs_b200 import camera as pycam  # noqa: E402
from vkgs_b200 import synth  # noqa: E402


def measure(r, cams, frames=40, warm=5):
    r.set_option(L.OPT_STAGE_TIMING, 1)
    acc, vis, pairs = {}, [], []
    for i in range(warm + frames):
        r.set_camera(block=cams[i % len(cams)])
        r.draw_device()
        if i >= warm:
            s = r.stats()
            for k in ("ms_cull", "ms_project", "ms_sort", "ms_bin", "ms_blend", "ms_total"):
                acc[k] = acc.get(k, 0.0) + s[k] / frames
            vis.append(s["visible_point_count"]); pairs.append(s["pair_count"])
            assert s["pair_overflow"] == 0
    r.set_option(L.OPT_STAGE_TIMING, 0)
    out = {k: round(v, 4) for k, v in acc.items()}
    out.update(fps=round(1e3 / acc["ms_total"], 1), visible_mean=float(np.mean(vis)), pairs_mean=float(np.mean(pairs)))
    return out


def orbit_cams(w, h, n, **kw):
    cams = []
    for i in range(n):
        c = pycam.orbit(w, h, theta_deg=30.0 + 360.0 * i / n, **kw)
        cams.append(vkgs_b200.camera_block(c.projection_matrix(), c.view_matrix(), c.eye()))
    return cams


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4", "c5"]
    n5 = int(sys.argv[sys.argv.index("--n5") + 1]) if "--n5" in sys.argv else 50_000_000
    if "c3" in which:
        w, h = 1920, 1080
        rows = synth.scene_garden()
        with vkgs_b200.Renderer(max_splats=rows.shape[0], max_width=w, max_height=h, max_pairs=128_000_000) as r:
            r.upload_splats(rows); del rows
            r.set_viewport(w, h)
            res = measure(r, orbit_cams(w, h, 16, r=12.0, phi_deg=60.0))
            print(json.dumps({"config": "C3 garden-shaped 5,834,734 splats, 1920x1080, zoomed out (r=12), 16-view orbit", **res}))
    if "c4" in which:
        w, h = 3840, 2160
        rows = synth.scene_bicycle()
        with vkgs_b200.Renderer(max_splats=rows.shape[0], max_width=w, max_height=h, max_pairs=128_000_000) as r:
            r.upload_splats(rows); del rows
            r.set_viewport(w, h)
            res = measure(r, orbit_cams(w, h, 45, r=1.5, phi_deg=70.0))   # one of 8 GPUs' share of a 360-view orbit
            print(json.dumps({"config": "C4 bicycle-shaped 6,131,954 splats, 3840x2160, 45 of 360 orbit views", **res}))
    if "c5" in which:
        w, h = 3840, 2160
        rows = synth.scene_large(n5)
        with vkgs_b200.Renderer(max_splats=n5, max_width=w, max_height=h, max_pairs=400_000_000) as r:
            r.upload_splats(rows); del rows
            r.set_viewport(w, h)
            cams = orbit_cams(w, h, 8, r=6.0, phi_deg=70.0)
            res = measure(r, cams, frames=16, warm=3)
            print(json.dumps({"config": f"C5 {n5:,} splats, 3840x2160, whole frame on one GPU", **res}))
            r.set_band(4 * h // 8, 5 * h // 8)
            res = measure(r, cams, frames=16, warm=3)
            print(json.dumps({"config": f"C5 {n5:,} splats, 3840x2160, band 5 of 8 (what one GPU of 8 runs: the cull drops what cannot reach the band)", **res}))


if __name__ == "__main__":
    main()
