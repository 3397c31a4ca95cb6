#!/usr/bin/env python
"""BASELINE.json configs[4]: a large scene rendered in screen-tile bands, one band per GPU, the scene replicated and the
finished bands gathered to rank 0 with NCCL (strong scaling of ONE frame; bench.py shards by view instead).
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/band_bench.py [--splats 50000000] [--frames 20]
Rank 0 prints one JSON line: ms per assembled frame (device time, max over ranks), the slowest band's stage times, and
the single-GPU whole-frame time of the same views for comparison (rank 0 renders it after the timed loop)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--splats", dest="n", type=int, default=50_000_000)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--equal", action="store_true", help="equal-height bands instead of load-balanced ones")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import vkgs_b200
    from vkgs_b200 import _lib as L
    from vkgs_b200 import camera as pycam
    from vkgs_b200 import dist as vdist
    from vkgs_b200 import synth

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, h = 1600, 900
    rows = synth.scene_large(args.n)
    r = vkgs_b200.Renderer(device=local, max_splats=args.n, max_width=w, max_height=h, max_pairs=256_000_000)
    r.upload_splats(rows)
    del rows
    r.set_viewport(w, h)
    cams = []
    for i in range(8):
        c = pycam.orbit(w, h, r=6.0, phi_deg=70.0, theta_deg=30.0 + 45.0 * i)
        cams.append(vkgs_b200.camera_block(c.projection_matrix(), c.view_matrix(), c.eye()))
    ap_balanced = "--equal" not in sys.argv
    if world > 1 and ap_balanced:
        # band edges balanced on the splat centres per row, summed over one whole frame per view of the orbit (every rank
        # holds the scene and computes the same edges).  One set of edges for the whole run: a band is part of the
        # recorded frame graph, changing it every frame would re-record the graph every frame.
        hist = np.zeros(h, np.float64)
        for cam in cams:
            r.set_camera(block=cam)
            r.draw_device()
            r.sync()
            hist += r.row_histogram()
        edges = vdist.balanced_band_edges(hist, world)
    else:
        edges = vdist.band_edges(h, world)
    edges_per_view = [edges] * len(cams)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    frame = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(frame) for _ in range(world)] if (world > 1 and rank == 0) else None

    def run(n_frames, banded):
        for i in range(n_frames):
            e = edges_per_view[i % len(cams)]
            band = (e[rank], e[rank + 1]) if banded else (0, 0)
            r.set_band(*band)
            r.set_camera(block=cams[i % len(cams)])
            r.draw_device(dst_ptr=frame.data_ptr(), stream=stream.cuda_stream)
            if world > 1 and banded:
                dist.gather(frame, gathered, dst=0)   # every rank's rows; rank 0 keeps rows [edges[g], edges[g+1]) of each

    band = world > 1
    run(args.warmup, band)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run(args.frames, band)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = vdist.max_over_ranks(e0.elapsed_time(e1) / args.frames, dev)
    # stage times of this rank's band
    r.set_option(L.OPT_STAGE_TIMING, 1)
    run(2, band)
    r.sync() if hasattr(r, "sync") else None
    st = r.stats()
    r.set_option(L.OPT_STAGE_TIMING, 0)
    stage_total = vdist.max_over_ranks(st["ms_total"], dev)
    whole = None
    if rank == 0:
        r.set_band(0, 0)
        for i in range(3):
            r.set_camera(block=cams[i % len(cams)]); r.draw_device(dst_ptr=frame.data_ptr(), stream=stream.cuda_stream)
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(8):
            r.set_camera(block=cams[i % len(cams)]); r.draw_device(dst_ptr=frame.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        whole = e0.elapsed_time(e1) / 8
        print(json.dumps({"config": f"C5 {args.n:,} splats, 1600x900, {world} band(s) on {world} GPU(s), bands gathered to rank 0 (NCCL)",
                          "band_edges": "balanced on the per-row splat histogram" if ap_balanced else "equal heights",
                          "edges_view0": edges_per_view[0],
                          "ms_per_frame": round(ms, 4), "fps": round(1e3 / ms, 1), "slowest_band_stages_ms_total": round(stage_total, 4),
                          "rank0_band": {k: round(st[k], 4) for k in ("ms_project", "ms_sort", "ms_bin", "ms_blend", "ms_total")},
                          "rank0_band_visible": st["visible_point_count"],
                          "whole_frame_one_gpu_ms": round(whole, 4), "speedup_vs_one_gpu": round(whole / ms, 2)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()


if __name__ == "__main__":
    main()
