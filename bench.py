#!/usr/bin/env python
"""Benchmarks of the frame path through the C ABI of libvkgsb.so.  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

  c2 (default, the headline; BASELINE.json configs[1]): frames/s at 1600x900 on the 6.1 M-splat SH3 'bicycle' scene.
      A step = one frame: camera block in, cull -> project -> sort -> bin -> blend, RGBA8 image out.  Views are sharded
      across ranks in contiguous blocks (weak scaling: K frames per rank), finished images delivered to rank 0.
      `value` is measured with the fp32 blend (VKGSB_BLEND_FP32); the reference-faithful VKGSB_BLEND_UNORM8 mode
      (B8G8R8A8_UNORM target, render_pass.cc:15) is timed the same way and reported beside it as `value_unorm8`,
      `stages_ms_unorm8`, `e2e_unorm8`.
  c4 (configs[3]): the same scene, 360-view orbit at 3840x2160 sharded by camera.
  c5 (configs[4]): 50 M splats at 3840x2160 in screen-tile bands, one band per GPU (strong scaling of one frame).
`--impl reference` times the CPU restatement of the reference's shaders (oracle/, OpenMP on all host cores; the
reference's Vulkan build is not runnable on this image: no loader, ICD or SDK, profiles/r02_host_vulkan_probe.txt) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "frames/s"
PARAM_BYTES = 560                 # sizeof(FrameParams): the per-frame host->device upload (a kernel argument)
E2E_BATCH = 32                    # views per vkgsb_draw_batch call in the end-to-end leg (an orbit is submitted in one
                                  # call when it has no more views: the call returns when every image is in host memory)
SMEM_BYTES_PER_ENTRY = 52         # blend stage: 3 x float4 raster record + 4-byte sub-tile mask per staged list entry
FP32_INSTR_PER_FRAGMENT = 20      # algorithmic FP32 instructions per fragment (the blend roofline's unit)

CONFIGS = {
    "c2": dict(metric="fps_1600x900_6.1M_splats_sh3", width=1600, height=900, n_splats=6_131_954, scene="bicycle",
               n_views=64, orbit=dict(r=1.5, phi_deg=70.0), theta0=30.0, max_pairs=64_000_000, scaling="weak",
               workload=("C2 bicycle-shaped 6,131,954 splats SH3, 1600x900, 64-view orbit r=1.5 phi=70deg "
                         "(~2 M visible, the reference's 'view 2' regime)")),
    "c4": dict(metric="views_per_s_3840x2160_6.1M_splats_orbit360", width=3840, height=2160, n_splats=6_131_954,
               scene="bicycle", n_views=360, orbit=dict(r=4.0, phi_deg=70.0), theta0=0.0, max_pairs=128_000_000,
               scaling="strong",
               workload="C4 bicycle-shaped 6,131,954 splats SH3, 3840x2160, 360-view orbit r=4 phi=70deg sharded by camera"),
    "c5": dict(metric="fps_3840x2160_50M_splats_bands", width=3840, height=2160, n_splats=50_000_000, scene="large",
               n_views=8, orbit=dict(r=6.0, phi_deg=70.0), theta0=30.0, max_pairs=400_000_000, scaling="strong",
               workload="C5 50,000,000 splats SH3 (70 % background), 3840x2160, camera r=6, one screen band per GPU"),
}
# kernels of one frame: set_params, cull classify / mixed, project, 3 depth onesweep passes, bin tiles / scan / place, blend
KERNELS_PER_FRAME = 11


def view_camera(cfg, i):
    from vkgs_b200 import camera as pycam
    cam = pycam.orbit(cfg["width"], cfg["height"], theta_deg=cfg["theta0"] + 360.0 * (i % cfg["n_views"]) / cfg["n_views"],
                      **cfg["orbit"])
    return cam.projection_matrix(), cam.view_matrix(), cam.eye()


def make_scene(cfg, dist=None, rank=0, world=1):
    """The configuration's synthetic scene.  The 50 M-row scene of c5 (12.4 GB) is generated once per node: rank 0 writes
    it to /dev/shm and the other ranks map it."""
    from vkgs_b200 import synth
    if cfg["scene"] == "bicycle":
        return synth.scene_bicycle(cfg["n_splats"])
    if world == 1 or dist is None:
        return synth.scene_large(cfg["n_splats"])
    token = [os.getpid()]
    dist.broadcast_object_list(token, src=0)
    path = f"/dev/shm/vkgsb_scene_{token[0]}.npy"
    if rank == 0:
        rows = synth.scene_large(cfg["n_splats"])
        np.save(path + ".tmp.npy", rows)
        os.rename(path + ".tmp.npy", path)
    else:
        while not os.path.exists(path):
            time.sleep(0.05)
        rows = np.load(path, mmap_mode="r")
    return rows


def drop_shared_scene(dist, rank, world):
    """After every rank has uploaded the scene: remove rank 0's copy in /dev/shm."""
    if world > 1:
        dist.barrier()
        if rank == 0:
            import glob
            for f in glob.glob(f"/dev/shm/vkgsb_scene_{os.getpid()}.npy*"):
                os.remove(f)


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
    """dram__bytes_read + dram__bytes_write of the projection stage's kernels per frame, from the newest committed
    `ncu --set full` capture (tools/ncu_traffic.py -> profiles/*_k_project_traffic.json); None when none is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_k_project_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        return float(d["dram_bytes_per_launch"]), os.path.basename(files[-1])
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------- the CPU arm
def cpu_sample(cfg, rows, view=0, state=None):
    """One step of the CPU arm: the CPU restatement of the reference path (oracle/, C + OpenMP on all host cores) on a
    bounded sample of one frame of the workload - the whole cull, sort and projection of the view, and the rasteriser
    (every fragment through the reference's UNORM8 blend, mode 1) on one 16-row band per host thread (4..16 bands) spread
    evenly over the image height, scaled to the full height.  Returns (frames/s estimate, seconds this step really took, description)."""
    from oracle import oracle as O
    from vkgs_b200 import synth
    O.use_all_cores()
    w, h = cfg["width"], cfg["height"]
    if state is None:
        state = {}
    if "scene" not in state:
        state["scene"] = O.activate(rows, synth.STANDARD_OFFSETS)
    scene = state["scene"]
    t0 = time.perf_counter()
    P, V, E = view_camera(cfg, view)
    cam = O.make_camera(P, V, E, w, h)
    keys, ids = O.cull(scene, O.compose_pvm(P, V))
    t1 = time.perf_counter()
    keys, ids = O.sort_pairs(keys, ids)
    t2 = time.perf_counter()
    inst = O.project(scene, ids, cam, 0)
    t3 = time.perf_counter()
    nb, band_rows = max(4, min(O.num_threads(), 16)), 16          # one tile band per host thread
    starts = [((h - band_rows) * (2 * b + 1) // (2 * nb)) // 16 * 16 for b in range(nb)]
    O.raster_band_list(inst, w, h, [y0 // 16 for y0 in starts], mode=1)     # one thread per band
    t4 = time.perf_counter()
    sampled = nb * band_rows
    frame_s = (t3 - t0) + (t4 - t3) * h / sampled
    desc = (f"1 frame of view {view}, V={len(ids)}: full cull {1e3*(t1-t0):.0f} ms + sort {1e3*(t2-t1):.0f} ms + projection "
            f"{1e3*(t3-t2):.0f} ms; rasteriser (every fragment, UNORM8 blend: the reference's ROP has no early exit) on {nb} "
            f"bands of {band_rows} rows at y={starts}, one thread per band, {1e3*(t4-t3):.0f} ms scaled x{h/sampled:.2f}")
    return 1.0 / frame_s, t4 - t0, desc, O.num_threads()


def run_reference(args, cfg):
    """The reference arm: K steps of the bounded CPU sample (a different orbit view each step); `value` is the frame rate
    the samples extrapolate to, `ms_per_step` what a step really took (so steps x ms_per_step is the run's wall time)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = make_scene(cfg)
    state, vals, secs, desc, cores = {}, [], [], "", 1
    steps = max(1, args.steps)
    warm = max(0, min(args.warmup, 2))
    budget_s = 240.0
    t_begin = time.perf_counter()
    done = 0
    for i in range(warm + steps):
        fps, s, desc, cores = cpu_sample(cfg, rows, view=(i * 7) % cfg["n_views"], state=state)
        if i >= warm:
            vals.append(fps); secs.append(s); done += 1
        if time.perf_counter() - t_begin > budget_s and done >= 1:
            break
    value = float(1.0 / np.mean(1.0 / np.asarray(vals)))      # frames / total estimated seconds
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": warm, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": desc + f"; {done} such steps (requested {steps}), frame_ms_estimate "
                                              f"{1e3/value:.0f}; the reference's own Vulkan path cannot run here "
                                              "(no Vulkan loader / lavapipe ICD / SDK on the box)",
                             "frame_ms_estimate": 1e3 / value},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- delivery to rank 0
class NcclGather:
    """Finished images go to rank 0 in batches of `every` frames; the gather of one batch (NCCL, its own stream)
    overlaps the rendering of the next into the other buffer."""
    name = "images gathered to rank 0 (NCCL, batches of 4 frames, double-buffered)"

    def __init__(self, torch, dist, dev, world, rank, h, w, every=4, nbuf=2):
        assert every == self.group
        self.dist, self.world, self.rank, self.every, self.nbuf = dist, world, rank, every, nbuf
        self.batches = [torch.empty((every, h, w, 4), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        self.gathered = [[torch.empty_like(self.batches[0]) for _ in range(world)] if (world > 1 and rank == 0) else None
                         for _ in range(nbuf)]
        self.pending = [None] * nbuf

    group = 4   # frames per vkgsb_draw_batch call = frames per gather

    def dst(self, i):
        """(device pointer of frame i, stride to frame i + 1) - i is a multiple of `group`"""
        b = (i // self.every) % self.nbuf
        if self.pending[b] is not None:   # the buffer's previous gather must have read it
            self.pending[b].wait()
            self.pending[b] = None
        return self.batches[b].data_ptr(), self.batches[b][0].numel()

    def sent(self, i, n):
        b = (i // self.every) % self.nbuf
        if self.world > 1:
            self.pending[b] = self.dist.gather(self.batches[b], self.gathered[b], dst=0, async_op=True)

    def drain(self):
        for b in range(self.nbuf):
            if self.pending[b] is not None:
                self.pending[b].wait()
                self.pending[b] = None


class PeerWrite:
    """Every rank renders straight into rank 0's memory: rank 0 allocates one buffer of `slots` x world images and
    exports it (CUDA IPC through the C ABI, vkgsb_shared_*); the other ranks map it and pass `their slot` as the frame's
    destination, so the blend kernel's pixel stores travel over NVLink and no kernel, copy or NCCL call runs on rank 0
    for the delivery.  A barrier every `slots` steps keeps a slot from being overwritten before rank 0 is done with it
    (never reached inside a timed region shorter than `slots` steps)."""
    name = "every rank's blend kernel writes its pixels into rank 0's buffer over NVLink (CUDA IPC peer mapping, no NCCL on the data path)"

    group = 4   # frames per vkgsb_draw_batch call

    def __init__(self, torch, dist, dev, world, rank, h, w, local, slots, images_per_slot=None):
        import vkgs_b200
        slots = (slots + self.group - 1) // self.group * self.group
        self.dist, self.world, self.rank, self.slots = dist, world, rank, slots
        self.img = h * w * 4
        self.V = vkgs_b200
        per_slot = world if images_per_slot is None else images_per_slot
        if rank == 0:
            self.base, handle = vkgs_b200.shared_create(local, slots * per_slot * self.img)
        else:
            self.base, handle = 0, b""
        obj = [handle]
        if world > 1:
            dist.broadcast_object_list(obj, src=0)
        if rank != 0:
            self.base = vkgs_b200.shared_open(local, obj[0])
        self.local = local

    def dst(self, i):
        """(device pointer of frame i, stride to frame i + 1) - i is a multiple of `group`"""
        if self.world > 1 and i > 0 and i % self.slots == 0:
            self.dist.barrier()
        return self.base + ((i % self.slots) * self.world + self.rank) * self.img, self.world * self.img

    def sent(self, i, n):
        pass

    def drain(self):
        pass

    def close(self):
        if self.rank == 0:
            self.V.shared_destroy(self.local, self.base)
        else:
            self.V.shared_close(self.local, self.base)


# ------------------------------------------------------------------------------------------------- the GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--blend", default="fp32", choices=["fp32", "unorm8"], help="blend mode of the headline `value`")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="how finished images reach rank 0 (N > 1)")
    ap.add_argument("--n-splats", type=int, default=0, help="override the scene size (experiments)")
    ap.add_argument("--band-cull", default="replicated", choices=["shared", "replicated"],
                    help="c5: the cull shared out over the ranks (vkgsb_group_*: each rank tests 1/N of the splats against "
                         "every band and writes the bands' bits into their GPUs) or repeated over the whole scene by every rank")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.n_splats:
        cfg["n_splats"] = args.n_splats
        cfg["workload"] += f" [scene size overridden: {args.n_splats}]"
    if args.impl == "reference":
        return run_reference(args, cfg)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import vkgs_b200
    from vkgs_b200 import _lib as L
    from vkgs_b200 import dist as vdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    W_, H_ = cfg["width"], cfg["height"]
    rows = make_scene(cfg, dist if world > 1 else None, rank, world)
    r = vkgs_b200.Renderer(device=local, max_splats=cfg["n_splats"], max_width=W_, max_height=H_, max_pairs=cfg["max_pairs"])
    if os.environ.get("VKGSB_SPATIAL_ORDER"):      # experiments: 0 keeps the file's order (no spatial order of the stored scene)
        r.set_option(L.OPT_SPATIAL_ORDER, int(os.environ["VKGSB_SPATIAL_ORDER"]))
    r.upload_splats(rows)
    if cfg["scene"] != "bicycle":
        drop_shared_scene(dist, rank, world)
    if not (world == 1 and not args.no_cpu_baseline and args.config == "c2"):
        del rows
    r.set_viewport(W_, H_)
    if os.environ.get("VKGSB_L2_PIN_MB"):          # experiments: how much of the splat centres is kept in L2
        r.set_option(L.OPT_L2_PIN_MB, int(os.environ["VKGSB_L2_PIN_MB"]))
    if os.environ.get("VKGSB_UNORM8_CUT_EXP"):
        r.set_option(L.OPT_UNORM8_CUT_EXP, int(os.environ["VKGSB_UNORM8_CUT_EXP"]))
    # a side stream: the legacy default stream's handle is 0, which the C ABI reads as "the renderer's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "c5":
        line = run_bands(args, cfg, r, torch, dist, vdist, L, dev, world, rank, local, stream, barrier)
    else:
        line = run_views(args, cfg, r, torch, dist, vdist, L, dev, world, rank, local, stream, barrier)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.config == "c2":
            fps, secs, desc, cores = cpu_sample(cfg, rows)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                    "frame_ms_estimate": 1e3 / fps, "sample_seconds": secs}
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_views(args, cfg, r, torch, dist, vdist, L, dev, world, rank, local, stream, barrier):
    """c2 / c4: views sharded across ranks, finished images delivered to rank 0."""
    import vkgs_b200
    K, W = args.steps, args.warmup
    W_, H_ = cfg["width"], cfg["height"]
    sptr = stream.cuda_stream
    img_bytes = W_ * H_ * 4
    if args.config == "c4":
        # strong scaling of one 360-view orbit: the views are dealt round-robin (dist.deal_views: neighbouring views cost
        # about the same, so every rank gets the same mix); --steps bounds the views per rank so that a default run
        # stays short
        mine = list(vdist.deal_views(cfg["n_views"], rank, world))
        K = min(K, len(mine))
        views = [mine[i % len(mine)] for i in range(max(K, W))]
        total_views = sum(min(args.steps, len(vdist.deal_views(cfg["n_views"], g, world))) for g in range(world))
    else:
        # weak scaling: K frames per rank, dealt round-robin like dist.deal_views: at step i rank g renders view i * world + g
        views = [i * world + rank for i in range(max(K, W))]
        total_views = world * K
    cams = [vkgs_b200.camera_block(*view_camera(cfg, v)) for v in views]
    if world > 1 and args.gather == "peer":
        slots = max(K, W) if max(K, W) * world * img_bytes <= (24 << 30) else max(8, (24 << 30) // (world * img_bytes))
        deliver = PeerWrite(torch, dist, dev, world, rank, H_, W_, local, slots)
    else:
        deliver = NcclGather(torch, dist, dev, world, rank, H_, W_)

    G = deliver.group

    def frames_device(k):
        """k frames through the public batch call, G views at a time (the orbit usage: consecutive frames run side by side
        on the device), every frame rendered straight into its destination"""
        for i in range(0, k, G):
            n = min(G, k - i)
            ptr, stride = deliver.dst(i)
            r.draw_batch([cams[(i + j) % len(cams)] for j in range(n)], dst_ptr=ptr, stride=stride, stream=sptr)
            deliver.sent(i, n)

    def timed(mode):
        r.set_blend_mode(mode)
        frames_device(W)
        deliver.drain()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        frames_device(W)   # the GPU stays under load while the sampler comes up
        deliver.drain()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.mark_begin()
        e0.record(stream)
        frames_device(K)
        deliver.drain()   # every image has arrived on rank 0 inside the timed region (peer writes: when the stream
        e1.record(stream)  # and the barrier below have completed)
        barrier()
        sampler.mark_end()
        ms = vdist.max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, sampler.stop()

    # ---- end to end through the public API with host buffers: camera blocks in, pixels out to pinned host memory,
    #      in orbit batches of E2E_BATCH views per vkgsb_draw_batch call (frame i crosses PCIe while frame i+1 renders;
    #      the call returns when every image of the batch is in host memory)
    host = torch.empty((E2E_BATCH, H_, W_, 4), dtype=torch.uint8).pin_memory()

    def e2e(mode):
        r.set_blend_mode(mode)

        def frames(k):
            done = 0
            while done < k:
                nb = min(E2E_BATCH, k - done)
                r.draw_batch_to_host_ptr([cams[(done + j) % len(cams)] for j in range(nb)], host.data_ptr(), stream=sptr)
                done += nb
        frames(E2E_BATCH)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        frames(K)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = vdist.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)), dev)  # device and host clocks
        return ms, int(host[0].sum().item())

    # ---- per-stage times (eager launches with events between stages) and the fragments the blend stage shades
    def stages(mode):
        r.set_blend_mode(mode)
        r.set_option(L.OPT_STAGE_TIMING, 1)
        acc = dict(ms_cull=0.0, ms_project=0.0, ms_sort=0.0, ms_bin=0.0, ms_blend=0.0, ms_total=0.0)
        vis, pairs, totals, retries = [], [], [], []
        for i in range(W + K):
            r.set_camera(block=cams[i % len(cams)])
            r.draw_device(stream=sptr)
            if i >= W:
                s = r.stats()
                for k in acc:
                    acc[k] += s[k]
                vis.append(s["visible_point_count"]); pairs.append(s["pair_count"]); totals.append(s["ms_total"])
                retries.append(s["blend_retries"])
        r.set_option(L.OPT_COUNT_FRAGMENTS, 1)
        frags = []
        for i in range(min(K, len(cams), 64)):
            r.set_camera(block=cams[i])
            r.draw_device(stream=sptr)
            frags.append(r.stats()["fragment_count"])
        r.set_option(L.OPT_COUNT_FRAGMENTS, 0)
        r.set_option(L.OPT_STAGE_TIMING, 0)
        st = {k[3:]: v / K for k, v in acc.items()}
        return dict(stages=st, V=float(np.mean(vis)), D=float(np.mean(pairs)), totals=totals, overflow=int(r.stats()["pair_overflow"]),
                    fragments=float(np.mean(frags)), retries=float(np.mean(retries)))

    FP32, U8 = L.BLEND_FP32, L.BLEND_UNORM8
    head, other = (FP32, U8) if args.blend == "fp32" else (U8, FP32)
    ms, clocks = timed(head)
    ms_o, clocks_o = timed(other) if args.config == "c2" else (None, None)
    e2e_ms, checksum = e2e(head)
    e2e_ms_o, _ = e2e(other) if args.config == "c2" else (None, None)
    S = stages(head)
    S_o = stages(other) if args.config == "c2" else None
    r.set_blend_mode(head)
    if hasattr(deliver, "close"):
        barrier()
        deliver.close()
    if rank != 0:
        return None

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    stage, V_mean = S["stages"], S["V"]
    alg_bytes = 12.0 * cfg["n_splats"] + 180.0 * V_mean            # reference-layout algorithmic bytes
    proj_gbs = alg_bytes / (stage["project"] * 1e-3) / 1e9
    dominant = max(("project", "sort", "bin", "blend"), key=lambda k: stage[k])
    traffic, traffic_src = ncu_traffic()
    frag_ceiling = 148 * 128 * sm_mhz * 1e6 / FP32_INSTR_PER_FRAGMENT     # FP32 lanes / 20 instructions

    def blend_block(S_):
        fps_ = S_["fragments"] / (S_["stages"]["blend"] * 1e-3)
        return {"fragments_per_frame": S_["fragments"], "fragments_per_s": fps_,
                "fp32_issue_frac": fps_ / frag_ceiling, "fp32_instr_per_fragment_assumed": FP32_INSTR_PER_FRAGMENT,
                "smem_bytes_per_entry": SMEM_BYTES_PER_ENTRY, "retries_per_frame": S_["retries"]}

    name = {FP32: "fp32", U8: "unorm8"}
    total = total_views if args.config == "c4" else world * K
    line = {
        "metric": cfg["metric"], "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "blend": name[head], "visible_mean": V_mean, "pairs_mean": S["D"],
                   "parallelism": (f"views dealt round-robin over {world} GPU(s), scene replicated, " + deliver.name)
                   if world > 1 else "single GPU",
                   "l2": f"inputs larger than L2: {144 * cfg['n_splats'] / 1e6:.0f} MB resident scene, its centres and visible payload lines streamed per frame, "
                         "a different camera each step",
                   "pair_overflow": S["overflow"]},
        "stages_ms": stage,
        # per-frame device time over the orbit (the views differ in cost): SURVEY.md 8(d) asks for the spread
        "frame_ms": {"p10": float(np.percentile(S["totals"], 10)), "p50": float(np.percentile(S["totals"], 50)),
                     "p90": float(np.percentile(S["totals"], 90))},
        "sort_gkeys_per_s": V_mean / (stage["sort"] * 1e-3) / 1e9,
        "sort_hbm_frac": 68.0 * V_mean / (stage["sort"] * 1e-3) / 1e9 / hbm_peak,
        "roofline": {"kernel": "k_cull_classify + k_cull_mixed + k_project (the projection stage)", "bound": "hbm", "achieved": proj_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": proj_gbs / hbm_peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_kind": peak_kind, "algorithmic_bytes": alg_bytes,
                     "share_of_step": stage["project"] / stage["total"], "dominant_by_time": dominant},
        "blend": blend_block(S),
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": PARAM_BYTES, "batch": E2E_BATCH,
                "d2h_bytes_per_step": img_bytes + 24, "checksum": checksum},
        "gpu_launches": KERNELS_PER_FRAME * K,
        "clocks": clocks,
    }
    if S_o is not None:
        o = name[other]
        line[f"value_{o}"] = total / (ms_o * 1e-3)
        line[f"ms_per_step_{o}"] = ms_o / K
        line[f"stages_ms_{o}"] = S_o["stages"]
        line[f"blend_{o}"] = blend_block(S_o)
        line[f"e2e_{o}"] = {"value": total / (e2e_ms_o * 1e-3), "unit": UNIT}
        line[f"clocks_{o}"] = clocks_o
    return line


def run_bands(args, cfg, r, torch, dist, vdist, L, dev, world, rank, local, stream, barrier):
    """c5: ONE frame per step, rank g renders band g (rows balanced on the per-row histogram of splat centres), the
    bands land in rank 0's frame.  value = assembled frames/s; strong scaling against the same frame on one GPU."""
    import vkgs_b200
    K, W = args.steps, args.warmup
    K = min(K, 40)
    W_, H_ = cfg["width"], cfg["height"]
    sptr = stream.cuda_stream
    cams = [vkgs_b200.camera_block(*view_camera(cfg, v)) for v in range(cfg["n_views"])]
    r.set_blend_mode(L.BLEND_UNORM8 if args.blend == "unorm8" else L.BLEND_FP32)
    if world > 1:
        # band edges balanced on the splat centres per row, summed over one whole frame per view of the orbit (every rank
        # holds the scene and computes the same edges).  One set of edges for the whole run: a band is part of the
        # recorded frame graph.
        hist = np.zeros(H_, np.float64)
        for cam in cams:
            r.set_camera(block=cam)
            r.draw_device()
            r.sync()
            hist += r.row_histogram()
        edges = vdist.balanced_band_edges(hist, world)
    else:
        edges = [0, H_]
    img_bytes = W_ * H_ * 4
    if world > 1 and args.gather == "peer":
        deliver = PeerWrite(torch, dist, dev, world, rank, H_, W_, local, 8, images_per_slot=1)  # ONE frame per slot: every rank writes its rows of it
    else:
        deliver = None
        frame = torch.zeros((H_, W_, 4), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(frame) for _ in range(world)] if (world > 1 and rank == 0) else None
    grouped = world > 1 and args.band_cull == "shared"
    if world > 1 and not grouped:
        # feedback: the centre histogram does not know what a band costs apart from its splats (the cull over its
        # frustum, the blend's per-pixel work), so the edges are re-cut twice from the bands' measured stage times
        for _ in range(2):
            r.set_band(edges[rank], edges[rank + 1])
            r.set_option(L.OPT_STAGE_TIMING, 1)
            t_band = 0.0
            for i in range(2 + 6):
                r.set_camera(block=cams[i % len(cams)])
                r.draw_device()
                if i >= 2:
                    t_band += r.stats()["ms_total"] / 6
            r.set_option(L.OPT_STAGE_TIMING, 0)
            r.sync()
            t_all = [None] * world
            dist.all_gather_object(t_all, float(t_band))
            edges = vdist.rebalance_band_edges(edges, t_all)
    if grouped:
        handles = [None] * world
        dist.all_gather_object(handles, r.group_export())
        r.group_join(rank, world, handles, edges)          # also sets this rank's band
    elif world > 1:
        r.set_band(edges[rank], edges[rank + 1])
    else:
        r.set_band(0, 0)

    def step(i):
        r.set_camera(block=cams[i % len(cams)])
        if deliver is not None:
            # every rank's blend kernel writes its band's rows straight into rank 0's frame (slot i % 8); the ranks meet
            # before the ring of slots wraps
            if i > 0 and i % 8 == 0:
                dist.barrier()
            r.draw_device(dst_ptr=deliver.base + (i % 8) * img_bytes, stream=sptr)
        else:
            r.draw_device(dst_ptr=frame.data_ptr(), stream=sptr)
            if world > 1:
                dist.gather(frame, gathered, dst=0)

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for i in range(K):
        step(i)
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = vdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop()
    r.set_option(L.OPT_STAGE_TIMING, 1)
    acc, vis = {}, []
    for i in range(3 + 8):
        r.set_camera(block=cams[i % len(cams)])
        r.draw_device(stream=sptr)
        if i >= 3:
            s = r.stats()
            for k in ("ms_cull", "ms_project", "ms_sort", "ms_bin", "ms_blend", "ms_total"):
                acc[k[3:]] = acc.get(k[3:], 0.0) + s[k] / 8
            vis.append(s["visible_point_count"])
    r.set_option(L.OPT_STAGE_TIMING, 0)
    slowest = vdist.max_over_ranks(acc["total"], dev)
    per_rank = [None] * world
    if world > 1:
        dist.all_gather_object(per_rank, dict(acc, visible=float(np.mean(vis))))
    if grouped:
        barrier()
        r.group_leave()
    if deliver is not None:
        barrier()
        deliver.close()
    if rank != 0:
        return None
    return {
        "metric": cfg["metric"], "value": K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "blend": args.blend, "band_edges": edges,
                   "parallelism": (f"{world} screen bands on {world} GPU(s), scene replicated, band edges balanced on the row "
                                   "histogram" + ("" if grouped else " and re-cut twice from the bands' measured times") + "; cull " + ("shared out over the ranks, each band's visibility bits written into its GPU "
                                                         "over NVLink (no collective); " if grouped else "repeated by every rank; ") + (PeerWrite.name if deliver is not None else "bands gathered to rank 0 (NCCL)"))
                   if world > 1 else "single GPU, whole frame",
                   "l2": f"inputs larger than L2: {144 * cfg['n_splats'] / 1e6:.0f} MB resident scene, its centres and visible payload lines streamed per frame",
                   "rank0_visible_mean": float(np.mean(vis))},
        "stages_ms": acc, "slowest_band_stages_ms_total": slowest,
        "bands": [{k: round(v, 4) for k, v in b.items()} for b in per_rank] if world > 1 else None,
        "gpu_launches": KERNELS_PER_FRAME * K, "clocks": clocks,
    }


if __name__ == "__main__":
    main()
