#!/usr/bin/env python
"""How much of a frame is not kernel time: device time per frame of the replayed graph at ONE view (CUDA events around
200 frames) next to the sum of the kernels' own durations from the ncu launch list of the same view
(tools/gpu_check.sh prints that sum).  usage: python tools/frame_overhead.py [view]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import vkgs_b200  # noqa: E402
from vkgs_b200 import synth  # noqa: E402

view = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rows = synth.scene_bicycle(bench.CONFIGS["c2"]["n_splats"])
r = vkgs_b200.Renderer(device=0, max_splats=bench.CONFIGS["c2"]["n_splats"], max_width=bench.CONFIGS["c2"]["width"], max_height=bench.CONFIGS["c2"]["height"], max_pairs=64_000_000)
r.upload_splats(rows)
del rows
r.set_viewport(bench.CONFIGS["c2"]["width"], bench.CONFIGS["c2"]["height"])
cam = vkgs_b200.camera_block(*bench.view_camera(bench.CONFIGS["c2"], view))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in ("graph", "eager+events"):
    r.set_option(vkgs_b200.OPT_STAGE_TIMING, 0 if mode == "graph" else 1)
    for _ in range(20):
        r.set_camera(block=cam)
        r.draw_device(stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(200):
        r.set_camera(block=cam)
        r.draw_device(stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"view {view} {mode}: {1e3 * e0.elapsed_time(e1) / 200:.1f} us per frame", r.stats() if mode != "graph" else "")
r.close()
