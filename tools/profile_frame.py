#!/usr/bin/env python
"""Driver for ncu: loads the bench scene (C2, BASELINE.json configs[1]) and draws a few frames with the kernels
launched eagerly, bracketed by cudaProfilerStart/Stop so `ncu --profile-from-start off` sees only steady-state frames.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 3
  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/frame \
      python tools/profile_frame.py --frames 1

Numbers printed under ncu are not bench values.
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--blend", default="fp32", choices=["fp32", "unorm8"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"])
    ap.add_argument("--view", type=int, default=0)
    args = ap.parse_args()

    import bench
    import vkgs_b200
    from vkgs_b200 import _lib as L
    from vkgs_b200 import synth

    rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
    if rt is None:
        for p in ("/usr/local/cuda/lib64/libcudart.so", "libcudart.so"):
            try:
                rt = ctypes.CDLL(p)
                break
            except OSError:
                pass
    if args.config == "c2":
        cfg = bench.CONFIGS["c2"]
        n, w, h = cfg["n_splats"], cfg["width"], cfg["height"]
        rows = synth.scene_bicycle(n)
        cam = vkgs_b200.camera_block(*bench.view_camera(cfg, args.view))
    else:
        from vkgs_b200 import camera as pycam
        n, w, h = 5_834_734, 1920, 1080
        rows = synth.scene_garden(n)
        c = pycam.orbit(w, h, r=12.0, phi_deg=60.0, theta_deg=45.0)  # zoomed out
        cam = vkgs_b200.camera_block(c.projection_matrix(), c.view_matrix(), c.eye())
    r = vkgs_b200.Renderer(device=0, max_splats=n, max_width=w, max_height=h, max_pairs=64_000_000)
    r.upload_splats(rows)
    del rows
    r.set_viewport(w, h)
    r.set_blend_mode(L.BLEND_UNORM8 if args.blend == "unorm8" else L.BLEND_FP32)
    r.set_option(L.OPT_STAGE_TIMING, 1)
    r.set_camera(block=cam)
    for _ in range(args.warm):
        r.draw_device()
    r.sync()
    if rt is not None:
        rt.cudaProfilerStart()
    for _ in range(args.frames):
        r.draw_device()
    r.sync()
    if rt is not None:
        rt.cudaProfilerStop()
    print(r.stats())
    r.close()


if __name__ == "__main__":
    main()
