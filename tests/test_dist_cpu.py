"""N>1 host logic on CPU: two `gloo` ranks shard an orbit by view and one frame by screen band (SURVEY.md §8e,
vkgs_b200/dist.py), gather to rank 0 and must reproduce the single-process result byte for byte.  The CPU oracle stands
in for the renderer here (tests may use it as the checker's renderer; the product never does): what is under test is
the sharding, the gather and the assembly, the very functions bench.py --gpus N and a band-sharded viewer call."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from vkgs_b200 import camera as pycam
from vkgs_b200 import dist as vdist
from vkgs_b200 import synth

W, H, N_VIEWS, N_SPLATS = 160, 96, 5, 3000


def _scene():
    return O.activate(synth.scene_c1(n=N_SPLATS, seed=9), synth.STANDARD_OFFSETS)


def _camera(i):
    c = pycam.orbit(W, H, r=2.5, phi_deg=60.0, theta_deg=360.0 * i / N_VIEWS)
    return c.projection_matrix(), c.view_matrix(), c.eye()


def _render_view(scene, i):
    P, V, E = _camera(i)
    return O.render(scene, O.make_camera(P, V, E, W, H), mode=0)["image"]


def _render_band(scene, i, y0, y1):
    P, V, E = _camera(i)
    cam = O.make_camera(P, V, E, W, H)
    keys, ids = O.cull(scene, O.compose_pvm(P, V))
    keys, ids = O.sort_pairs(keys, ids)
    inst = O.project(scene, ids, cam, 0)
    img = O.raster_rows(inst, W, H, y0, y1, mode=0)
    img[:y0] = 0
    img[y1:] = 0   # rows outside the band are not this rank's to deliver
    return img


def _row_histogram(scene, i):
    """What vkgsb_row_histogram returns on the GPU: visible splat centres per image row."""
    P, V, E = _camera(i)
    keys, ids = O.cull(scene, O.compose_pvm(P, V))
    inst = O.project(scene, ids, O.make_camera(P, V, E, W, H), 0)
    cpy = np.float32(H / 2) * inst[:, 1] + np.float32(H / 2 - 0.5)
    cpy = cpy[~np.isnan(cpy)]
    return np.bincount(np.clip(np.rint(cpy), 0, H - 1).astype(np.int64), minlength=H)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = _scene()
        # ---- by view: contiguous camera blocks, padded to the largest block so the gather is rectangular
        mine = vdist.shard_views(N_VIEWS, rank, world)
        kmax = max(len(vdist.shard_views(N_VIEWS, r, world)) for r in range(world))
        local = torch.zeros((kmax, H, W, 4), dtype=torch.uint8)
        for j, i in enumerate(mine):
            local[j] = torch.from_numpy(_render_view(scene, i))
        got = vdist.gather_images(local, dst=0)
        # ---- by band: one view, rank g owns rows [edges[g], edges[g+1])
        edges = vdist.band_edges(H, world)
        band = torch.from_numpy(_render_band(scene, 1, edges[rank], edges[rank + 1]))[None]
        got_b = vdist.gather_images(band, dst=0)
        # ---- by load-balanced band: edges from the per-row histogram of splat centres (every rank computes the same)
        load = _row_histogram(scene, 1).astype(np.float64)
        load[: H // 3] *= 6.0   # as if the upper third were the expensive part: the edge moves off the tile grid
        bal = vdist.balanced_band_edges(load, world, min_rows=4)
        band2 = torch.from_numpy(_render_band(scene, 1, bal[rank], bal[rank + 1]))[None]
        got_c = vdist.gather_images(band2, dst=0)
        slowest = vdist.max_over_ranks(float(rank + 1))
        if rank == 0:
            views = vdist.assemble_views(got, N_VIEWS).numpy()
            frame = vdist.assemble_bands(got_b, edges).numpy()[0]
            frame2 = vdist.assemble_bands(got_c, bal).numpy()[0]
            np.savez(out_path, views=views, frame=frame, frame2=frame2, slowest=slowest, edges=np.array(edges), bal=np.array(bal))
        else:
            assert got is None and got_b is None and got_c is None
    finally:
        dist.destroy_process_group()


def test_shard_views_partitions_the_orbit():
    for n in (0, 1, 5, 64, 360):
        for world in (1, 2, 3, 4, 8):
            blocks = [vdist.shard_views(n, r, world) for r in range(world)]
            assert [i for b in blocks for i in b] == list(range(n))
            assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1


def test_band_edges_cover_the_frame_tile_aligned():
    for h in (16, 96, 600, 900, 1080, 2160):
        for world in (1, 2, 4, 8):
            e = vdist.band_edges(h, world)
            assert e[0] == 0 and e[-1] == h and len(e) == world + 1
            assert all(a <= b for a, b in zip(e, e[1:]))
            assert all(x % 16 == 0 for x in e[:-1])


@pytest.mark.timeout(300)
def test_two_gloo_ranks_reproduce_the_single_process_result(tmp_path):
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    scene = _scene()
    want = np.stack([_render_view(scene, i) for i in range(N_VIEWS)])
    assert np.array_equal(z["views"], want), "view-sharded orbit differs from the single-process orbit"
    assert np.array_equal(z["frame"], want[1]), "bands do not concatenate to the full frame"
    assert float(z["slowest"]) == 2.0  # max over ranks, the timing rule of bench.py
    assert list(z["edges"]) == vdist.band_edges(H, 2)
    assert np.array_equal(z["frame2"], want[1]), "load-balanced bands do not concatenate to the full frame"
    bal = list(z["bal"])
    assert bal[0] == 0 and bal[-1] == H and bal != vdist.band_edges(H, 2)


def test_balanced_band_edges_cover_the_frame_and_even_out_the_load():
    rng = np.random.default_rng(3)
    h = 900
    load = rng.integers(0, 50, h).astype(np.float64)
    load[380:520] += 4000            # the scene sits in the middle rows
    for world in (1, 2, 4, 8):
        e = vdist.balanced_band_edges(load, world)
        assert e[0] == 0 and e[-1] == h and len(e) == world + 1
        assert all(b - a >= 8 for a, b in zip(e[:-1], e[1:]))
        per = [load[a:b].sum() for a, b in zip(e[:-1], e[1:])]
        equal = [load[a:b].sum() for a, b in zip(vdist.band_edges(h, world)[:-1], vdist.band_edges(h, world)[1:])]
        assert max(per) <= max(equal) + 1e-9
        if world == 8:
            assert max(per) < 0.25 * load.sum() < max(equal)      # equal-height bands leave most of it to two ranks
    assert vdist.balanced_band_edges(np.zeros(64), 4) == [0, 16, 32, 48, 64]


def test_rebalance_band_edges_moves_rows_from_slow_bands_to_fast_ones():
    """Feedback step of the band partition: with a cost model of a constant per band plus a share per row weight, a few
    steps even the band times out; the bands always cover the frame and keep their minimum height."""
    rng = np.random.default_rng(5)
    h, world = 2160, 8
    w = np.exp(-0.5 * ((np.arange(h) - 1100) / 250.0) ** 2) + 0.02          # row weights: most of the load mid-frame
    cum = np.concatenate([[0.0], np.cumsum(w)])

    def times(e):
        return [0.45 + 8.0 * (cum[e[g + 1]] - cum[e[g]]) / cum[-1] + 2e-4 * (e[g + 1] - e[g]) for g in range(world)]

    e = [h * g // world for g in range(world + 1)]                            # equal rows: the middle bands carry the load
    spread0 = max(times(e)) - min(times(e))
    for _ in range(4):
        e = vdist.rebalance_band_edges(e, times(e))
        assert e[0] == 0 and e[-1] == h and all(b - a >= 8 for a, b in zip(e, e[1:]))
    t = times(e)
    assert max(t) - min(t) < 0.25 * spread0 and max(t) < 1.05 * np.mean(t)
    assert vdist.rebalance_band_edges([0, 10, 20], [0.0, 0.0]) == [0, 10, 20]   # nothing measured: unchanged
    assert vdist.rebalance_band_edges([0, 32, 64], [1.0, float("nan")]) == [0, 32, 64]


def _scene_worker(rank, world, port, out):
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = dict(bench.CONFIGS["c5"])
        cfg["n_splats"] = 120_000
        rows = bench.make_scene(cfg, dist, rank, world)          # rank 0 generates, rank 1 maps rank 0's copy
        np.save(f"{out}.{rank}.npy", np.asarray(rows))
        kind = type(rows).__name__
        bench.drop_shared_scene(dist, rank, world)
        dist.barrier()
        import glob
        assert not glob.glob("/dev/shm/vkgsb_scene_*") or rank != 0
        with open(f"{out}.{rank}.txt", "w") as f:
            f.write(kind)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_c5_scene_is_generated_once_per_node_and_is_thread_count_independent(tmp_path):
    out = str(tmp_path / "scene")
    mp.spawn(_scene_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    want = synth.scene_large(120_000, threads=1)
    assert np.array_equal(synth.scene_large(120_000, chunk=1_000_000, threads=3), want)
    for rank in range(2):
        assert np.array_equal(np.load(f"{out}.{rank}.npy"), want)
    assert open(f"{out}.1.txt").read() == "memmap"
