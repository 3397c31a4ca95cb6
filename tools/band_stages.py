#!/usr/bin/env python
"""C5 on ONE GPU: the stage times of the whole frame and of each of the W balanced screen bands a W-GPU run would give
its ranks (bench.py --config c5 --band-cull replicated), one after the other on the same renderer.  The slowest band is
what one assembled frame costs on W GPUs.
  python tools/band_stages.py [--n 50000000] [--bands 8] [--frames 12]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50_000_000)
    ap.add_argument("--bands", type=int, default=8)
    ap.add_argument("--frames", type=int, default=12)
    a = ap.parse_args()
    import bench
    import vkgs_b200
    from vkgs_b200 import _lib as L
    from vkgs_b200 import dist as vdist
    cfg = dict(bench.CONFIGS["c5"])
    cfg["n_splats"] = a.n
    W_, H_ = cfg["width"], cfg["height"]
    rows = bench.make_scene(cfg)
    r = vkgs_b200.Renderer(max_splats=a.n, max_width=W_, max_height=H_, max_pairs=cfg["max_pairs"])
    r.upload_splats(rows)
    del rows
    r.set_viewport(W_, H_)
    cams = [vkgs_b200.camera_block(*bench.view_camera(cfg, v)) for v in range(cfg["n_views"])]
    hist = np.zeros(H_, np.float64)
    for cam in cams:
        r.set_camera(block=cam)
        r.draw_device()
        r.sync()
        hist += r.row_histogram()
    edges = vdist.balanced_band_edges(hist, a.bands)

    def measure(y0, y1):
        r.set_band(y0, y1)
        r.set_option(L.OPT_STAGE_TIMING, 1)
        acc, vis = {}, []
        for i in range(3 + a.frames):
            r.set_camera(block=cams[i % len(cams)])
            r.draw_device()
            if i >= 3:
                s = r.stats()
                for k in ("ms_cull", "ms_project", "ms_sort", "ms_bin", "ms_blend", "ms_total"):
                    acc[k] = acc.get(k, 0.0) + s[k] / a.frames
                vis.append(s["visible_point_count"])
        r.set_option(L.OPT_STAGE_TIMING, 0)
        out = {k[3:]: round(v, 4) for k, v in acc.items()}
        out["visible"] = float(np.mean(vis))
        return out

    whole = measure(0, 0)
    print(json.dumps({"what": f"C5 {a.n:,} splats whole frame", **whole}))
    worst = 0.0
    for g in range(a.bands):
        b = measure(edges[g], edges[g + 1])
        worst = max(worst, b["total"])
        print(json.dumps({"what": f"band {g} rows [{edges[g]}, {edges[g + 1]})", **b}))
    print(json.dumps({"what": "summary", "whole_ms": whole["total"], "slowest_band_ms": round(worst, 4),
                      "ratio": round(whole["total"] / worst, 2), "edges": edges}))
    r.close()


if __name__ == "__main__":
    main()
