#!/usr/bin/env python
"""Headline benchmark: frames/s at 1600x900 on the 6.1 M-splat SH3 synthetic 'bicycle' scene (BASELINE.json
configs[1]) through the C ABI of libvkgsb.so.  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one frame: camera block in, project -> sort -> bin -> blend, RGBA8 image out.  Views are sharded across
ranks (weak scaling: K frames per rank, different cameras), images gathered to rank 0 with NCCL.
`--impl reference` times the CPU restatement of the reference's shaders (oracle/, OpenMP on all host cores; the
reference's Vulkan build is not runnable here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fps_1600x900_6.1M_splats_sh3"
UNIT = "frames/s"
WIDTH, HEIGHT = 1600, 900
N_SPLATS = 6_131_954
N_VIEWS = 64                      # orbit the steps cycle through
ORBIT = dict(r=1.5, phi_deg=70.0)  # ~2 M visible of 6.1 M: the reference's "view 2" regime (DETAILS.md:72)
E2E_BATCH = 8                     # views per vkgsb_draw_batch call in the end-to-end leg
PARAM_BYTES = 548                 # sizeof(FrameParams): the per-frame host->device upload (a kernel argument)
WORKLOAD = ("C2 bicycle-shaped 6,131,954 splats SH3, 1600x900, 64-view orbit r=1.5 phi=70deg "
            "(~2 M visible, the reference's 'view 2' regime)")
KERNELS_PER_FRAME = 9             # set_params, project, 3 depth onesweep passes, bin count / scan / place, blend


def view_camera(i):
    from vkgs_b200 import camera as pycam
    cam = pycam.orbit(WIDTH, HEIGHT, r=ORBIT["r"], phi_deg=ORBIT["phi_deg"], theta_deg=30.0 + 360.0 * (i % N_VIEWS) / N_VIEWS)
    return cam.projection_matrix(), cam.view_matrix(), cam.eye()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe).  nvidia-smi takes a few
    hundred ms to deliver its first line and the timed region is shorter than that, so the sampler is started before the
    warm-up, start() returns once the first sample is in, and stop() keeps the samples stamped inside
    [mark_begin(), mark_end()] (the nearest ones around it when the region fell between two samples)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None
        self.t0 = self.t1 = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
            deadline = time.time() + 3.0
            while time.time() < deadline and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.05)  # one more sample behind the region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 10:
                    continue
                try:
                    rows.append((self._epoch(p[0]), float(p[2]), float(p[3]), p[6:10]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t0 <= r[0] <= self.t1]
        if not inside and rows and self.t0 is not None:   # the region fell between two 20 ms samples: its neighbours
            mid = 0.5 * (self.t0 + self.t1)
            inside = sorted((r for r in rows if r[0] is not None), key=lambda r: abs(r[0] - mid))[:2]
        if not inside:
            inside = rows
        reasons = set()
        for _, _, _, flags in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out["sm_mhz"] = float(np.median([r[1] for r in inside]))
            out["sm_max_mhz"] = float(max(r[2] for r in inside))
            out["samples"] = len(inside)
        out["reasons"] = sorted(reasons)
        return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic():
    """dram__bytes_read + dram__bytes_write per k_project launch, from the newest committed `ncu --set full` capture
    (tools/ncu_traffic.py -> profiles/*_k_project_traffic.json); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_k_project_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        return float(d["dram_bytes_per_launch"]), os.path.basename(files[-1])
    except Exception:
        return None, None


def cpu_sample(rows_fn, threads_note=""):
    """CPU restatement of the reference path (oracle, OpenMP) on a bounded sample: the whole cull / sort / projection
    of one frame, and the rasteriser on a 16-row band scaled to the frame height."""
    from oracle import oracle as O
    from vkgs_b200 import synth
    O.use_all_cores()
    rows = rows_fn()
    t0 = time.perf_counter()
    scene = O.activate(rows, synth.STANDARD_OFFSETS)
    del rows
    P, V, E = view_camera(0)
    cam = O.make_camera(P, V, E, WIDTH, HEIGHT)
    t1 = time.perf_counter()
    keys, ids = O.cull(scene, O.compose_pvm(P, V))
    t2 = time.perf_counter()
    keys, ids = O.sort_pairs(keys, ids)
    t3 = time.perf_counter()
    inst = O.project(scene, ids, cam, 0)
    t4 = time.perf_counter()
    band = (HEIGHT // 2 // 16) * 16
    rows_sampled = 16 * max(1, O.num_threads())           # one tile band per thread
    rows_sampled = min(rows_sampled, 128)
    r0 = max(0, band - rows_sampled // 2)
    O.raster_rows(inst, WIDTH, HEIGHT, r0, r0 + rows_sampled, mode=0)
    t5 = time.perf_counter()
    raster_full = (t5 - t4) * HEIGHT / rows_sampled
    frame_s = (t2 - t1) + (t3 - t2) + (t4 - t3) + raster_full
    return dict(value=1.0 / frame_s, unit=UNIT, cores=O.num_threads(), kind="port",
                sample=(f"1 frame, view 0, V={len(ids)}: full cull {1e3*(t2-t1):.0f} ms + sort {1e3*(t3-t2):.0f} ms + "
                        f"projection {1e3*(t4-t3):.0f} ms; rasteriser on rows [{r0},{r0+rows_sampled}) "
                        f"{1e3*(t5-t4):.0f} ms scaled x{HEIGHT/rows_sampled:.1f} (mid-frame band)"),
                frame_ms_estimate=1e3 * frame_s), len(ids)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vkgs_b200 import synth
    rows_holder = {}

    def rows_fn():
        if "r" not in rows_holder:
            rows_holder["r"] = synth.scene_bicycle(N_SPLATS)
        return rows_holder["r"]

    vals = []
    cb = None
    for i in range(max(1, min(args.steps, 3)) + min(args.warmup, 1)):
        cb, v = cpu_sample(rows_fn)
        vals.append(cb["value"])
    value = float(np.mean(vals[min(args.warmup, 1):]))
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_view": 0,
                       "note": "CPU restatement of the reference's shaders (oracle/, C + OpenMP); the reference's Vulkan "
                               "build / lavapipe is not available on this image"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--blend", default="fp32", choices=["fp32", "unorm8"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import vkgs_b200
    from vkgs_b200 import _lib as L
    from vkgs_b200 import dist as vdist
    from vkgs_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, args.warmup
    rows = synth.scene_bicycle(N_SPLATS)
    r = vkgs_b200.Renderer(device=local, max_splats=N_SPLATS, max_width=WIDTH, max_height=HEIGHT, max_pairs=64_000_000)
    r.upload_splats(rows)
    del rows
    r.set_viewport(WIDTH, HEIGHT)
    r.set_blend_mode(L.BLEND_UNORM8 if args.blend == "unorm8" else L.BLEND_FP32)
    # a side stream: the legacy default stream's handle is 0, which the C ABI reads as "the renderer's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    # the orbit's views are dealt round-robin: at step i rank g renders view i * world + g (neighbours on the orbit)
    cams = [vkgs_b200.camera_block(*view_camera(i * world + rank)) for i in range(max(K, W))]
    img_bytes = WIDTH * HEIGHT * 4
    GATHER_EVERY, NBUF = 4, 2
    # finished images go to rank 0 in batches of GATHER_EVERY frames; the gather of one batch (NCCL, its own stream)
    # overlaps the rendering of the next into the other buffer
    batches = [torch.empty((GATHER_EVERY, HEIGHT, WIDTH, 4), dtype=torch.uint8, device=dev) for _ in range(NBUF)]
    gathered = [[torch.empty_like(batches[0]) for _ in range(world)] if (world > 1 and rank == 0) else None
                for _ in range(NBUF)]
    pending = [None] * NBUF

    def drain():
        for b in range(NBUF):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    def frame_device(i):
        b, j = (i // GATHER_EVERY) % NBUF, i % GATHER_EVERY
        if j == 0 and pending[b] is not None:   # the buffer's previous gather must have read it
            pending[b].wait()
            pending[b] = None
        r.set_camera(block=cams[i % len(cams)])
        r.draw_device(dst_ptr=batches[b][j].data_ptr(), stream=sptr)
        if world > 1 and j == GATHER_EVERY - 1:
            pending[b] = dist.gather(batches[b], gathered[b], dst=0, async_op=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    for i in range(W):
        frame_device(i)
    drain()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):   # the GPU stays under load while the sampler comes up
        frame_device(i)
    drain()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for i in range(K):
        frame_device(i)
    drain()   # every gathered image has arrived on rank 0 inside the timed region
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = vdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop()
    st = r.stats()

    # ---- end to end through the public API with host buffers (`e2e`): camera blocks in, pixels out to pinned host
    #      memory, in orbit batches of E2E_BATCH views per vkgsb_draw_batch call (the C4 usage pattern: frame i crosses
    #      PCIe while frame i+1 renders; the call returns when every image of the batch is in host memory)
    host = torch.empty((E2E_BATCH, HEIGHT, WIDTH, 4), dtype=torch.uint8).pin_memory()

    def e2e_frames(k0, k):
        done = 0
        while done < k:
            nb = min(E2E_BATCH, k - done)
            r.draw_batch_to_host_ptr([cams[(k0 + done + j) % len(cams)] for j in range(nb)], host.data_ptr(), stream=sptr)
            done += nb

    e2e_frames(0, E2E_BATCH)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    e2e_frames(0, K)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = vdist.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)), dev)  # device and host clocks
    checksum = int(host[0].sum().item())

    # ---- per-stage times (eager launches with events between stages)
    r.set_option(L.OPT_STAGE_TIMING, 1)
    acc = dict(ms_project=0.0, ms_sort=0.0, ms_bin=0.0, ms_blend=0.0, ms_total=0.0)
    vis, pairs, totals = [], [], []
    for i in range(W + K):
        r.set_camera(block=cams[i % len(cams)])
        r.draw_device(stream=sptr)
        if i >= W:
            s = r.stats()
            for k in acc:
                acc[k] += s[k]
            vis.append(s["visible_point_count"]); pairs.append(s["pair_count"]); totals.append(s["ms_total"])
    r.set_option(L.OPT_STAGE_TIMING, 0)
    stage = {k: v / K for k, v in acc.items()}
    V_mean, D_mean = float(np.mean(vis)), float(np.mean(pairs))

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 12.0 * N_SPLATS + 180.0 * V_mean            # reference-layout algorithmic bytes
    proj_gbs = alg_bytes / (stage["ms_project"] * 1e-3) / 1e9
    sort_gkeys = V_mean / (stage["ms_sort"] * 1e-3) / 1e9
    dominant = max(("ms_project", "ms_sort", "ms_bin", "ms_blend"), key=lambda k: stage[k])
    traffic, traffic_src = ncu_traffic()

    if rank == 0:
        fps = world * K / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "blend": args.blend, "visible_mean": V_mean, "pairs_mean": D_mean,
                       "parallelism": f"views sharded over {world} GPU(s), scene replicated, images gathered to rank 0 (NCCL)"
                       if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: 834 MB resident scene streamed per frame, a different camera each step",
                       "pair_overflow": int(st["pair_overflow"])},
            "stages_ms": {"project": stage["ms_project"], "sort": stage["ms_sort"], "bin": stage["ms_bin"],
                          "blend": stage["ms_blend"], "total": stage["ms_total"]},
            # per-frame device time over the orbit (the views differ in cost): SURVEY.md 8(d) asks for the spread
            "frame_ms": {"p10": float(np.percentile(totals, 10)), "p50": float(np.percentile(totals, 50)),
                         "p90": float(np.percentile(totals, 90))},
            "sort_gkeys_per_s": sort_gkeys,
            "sort_hbm_frac": 68.0 * V_mean / (stage["ms_sort"] * 1e-3) / 1e9 / hbm_peak,
            "roofline": {"kernel": "k_project", "bound": "hbm", "achieved": proj_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": proj_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_kind": peak_kind,
                         "algorithmic_bytes": alg_bytes, "share_of_step": stage["ms_project"] / stage["ms_total"],
                         "dominant_by_time": dominant.replace("ms_", "")},
            "e2e": {"value": world * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": PARAM_BYTES, "batch": E2E_BATCH,
                    "d2h_bytes_per_step": img_bytes + 12, "checksum": checksum},
            "gpu_launches": KERNELS_PER_FRAME * K,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_sample(lambda: synth.scene_bicycle(N_SPLATS))
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
