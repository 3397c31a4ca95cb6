"""Seeded synthetic 3DGS scenes in the standard 62-float PLY row layout.

The reference has no fixtures; these are the inputs SURVEY.md §8(d) /
BASELINE.md §3 define so that the reference loader (src/vkgs/engine/splat_load_thread.cc:55-135)
could read the very same file.  Property order is the one the 3DGS trainer writes:
x,y,z,nx,ny,nz,f_dc_0..2,f_rest_0..44,opacity,scale_0..2,rot_0..3 (all `float`, binary LE).
"""
from __future__ import annotations

import numpy as np

PLY_PROPS = (
    ["x", "y", "z", "nx", "ny", "nz"]
    + [f"f_dc_{i}" for i in range(3)]
    + [f"f_rest_{i}" for i in range(45)]
    + ["opacity"]
    + [f"scale_{i}" for i in range(3)]
    + [f"rot_{i}" for i in range(4)]
)
ROW_FLOATS = len(PLY_PROPS)  # 62
assert ROW_FLOATS == 62


def offsets_from_props(props) -> np.ndarray:
    """The 60-entry float-offset table of splat_load_thread.cc:114-135.

    0-2 xyz, 3-5 scale, 6-9 rot_1,rot_2,rot_3,rot_0 (x,y,z,w), 10+16c+k SH (k=0 f_dc_c, k>=1
    f_rest_{15c+k-1}), 58 opacity, 59 row stride in floats.
    """
    pos = {p: i for i, p in enumerate(props)}
    off = np.zeros(60, dtype=np.uint32)
    off[0:3] = [pos["x"], pos["y"], pos["z"]]
    off[3:6] = [pos["scale_0"], pos["scale_1"], pos["scale_2"]]
    off[6:10] = [pos["rot_1"], pos["rot_2"], pos["rot_3"], pos["rot_0"]]
    for c in range(3):
        off[10 + 16 * c] = pos[f"f_dc_{c}"]
        for k in range(15):
            off[10 + 16 * c + 1 + k] = pos[f"f_rest_{15 * c + k}"]
    off[58] = pos["opacity"]
    off[59] = len(props)
    return off


STANDARD_OFFSETS = offsets_from_props(PLY_PROPS)

_COL = {p: i for i, p in enumerate(PLY_PROPS)}
_REST_BAND = np.array([1] * 3 + [2] * 5 + [3] * 7)  # SH band of f_rest index k (per channel)


def _fill_common(rows: np.ndarray, rng: np.random.Generator, dc_sigma=0.8, rest_sigma=0.15):
    n = rows.shape[0]
    rows[:, _COL["nx"]:_COL["nz"] + 1] = 0.0
    rows[:, _COL["f_dc_0"]:_COL["f_dc_2"] + 1] = rng.standard_normal((n, 3), dtype=np.float32) * dc_sigma
    sig = np.tile(rest_sigma / _REST_BAND, 3).astype(np.float32)  # channel-major, 15 per channel
    r0 = _COL["f_rest_0"]
    rows[:, r0:r0 + 45] = rng.standard_normal((n, 45), dtype=np.float32) * sig
    q = rng.standard_normal((n, 4), dtype=np.float32)
    q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    rows[:, _COL["rot_0"]:_COL["rot_3"] + 1] = q


def scene_c1(n: int = 100_000, seed: int = 1001) -> np.ndarray:
    """C1: gaussian ball, reference default camera sees all of it."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, ROW_FLOATS), dtype=np.float32)
    _fill_common(rows, rng)
    p = rng.standard_normal((n, 3), dtype=np.float32) * 0.6
    r = np.linalg.norm(p, axis=1, keepdims=True)
    p *= np.minimum(1.0, 1.999 / np.maximum(r, 1e-12)).astype(np.float32)
    rows[:, 0:3] = p
    ls = rng.standard_normal((n, 1), dtype=np.float32) * 0.8 - 4.0
    rows[:, _COL["scale_0"]:_COL["scale_2"] + 1] = ls + rng.standard_normal((n, 3), dtype=np.float32) * 0.5
    rows[:, _COL["opacity"]] = rng.standard_normal(n, dtype=np.float32) * 2.5 + 0.5
    return rows


def _blob_mixture(rng, n, nblob, extent, blob_sigma):
    centers = (rng.random((nblob, 3), dtype=np.float32) - 0.5) * np.asarray(extent, dtype=np.float32)
    axes = np.exp(rng.standard_normal((nblob, 3), dtype=np.float32) * 0.5) * blob_sigma
    which = rng.integers(0, nblob, size=n)
    return centers[which] + rng.standard_normal((n, 3), dtype=np.float32) * axes[which]


def _bimodal_opacity(rng, n):
    hi = rng.random(n) < 0.6
    o = np.where(hi, rng.standard_normal(n) * 1.0 + 3.0, rng.standard_normal(n) * 1.5 - 2.0)
    return o.astype(np.float32)


def scene_bicycle(n: int = 6_131_954, seed: int = 2002, background: float = 0.45) -> np.ndarray:
    """C2 'bicycle-shaped': foreground blob mixture in a 4x2x4 slab + log-uniform background shell."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, ROW_FLOATS), dtype=np.float32)
    _fill_common(rows, rng)
    nbg = int(n * background)
    nfg = n - nbg
    is_bg = np.zeros(n, dtype=bool)
    is_bg[rng.permutation(n)[:nbg]] = True  # interleave: ids carry no spatial order

    fg = _blob_mixture(rng, nfg, 64, (4.0, 2.0, 4.0), 0.35)
    rows[~is_bg, 0:3] = fg
    ls = rng.standard_normal((nfg, 1), dtype=np.float32) * 1.1 - 4.6
    rows[~is_bg, _COL["scale_0"]:_COL["scale_2"] + 1] = ls + rng.standard_normal((nfg, 3), dtype=np.float32) * 0.5

    rad = np.exp(rng.uniform(np.log(4.0), np.log(40.0), nbg)).astype(np.float32)
    d = rng.standard_normal((nbg, 3), dtype=np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rows[is_bg, 0:3] = d * rad[:, None]
    ls = rng.standard_normal((nbg, 1), dtype=np.float32) * 1.0 - 2.5 + np.log(rad / 40.0)[:, None]
    rows[is_bg, _COL["scale_0"]:_COL["scale_2"] + 1] = ls + rng.standard_normal((nbg, 3), dtype=np.float32) * 0.5
    rows[:, _COL["opacity"]] = _bimodal_opacity(rng, n)
    return rows


def scene_garden(n: int = 5_834_734, seed: int = 3003) -> np.ndarray:
    """C3 'garden-shaped': ground disc (40 %), central table object (25 %), hedge ring (35 %)."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, ROW_FLOATS), dtype=np.float32)
    _fill_common(rows, rng)
    kind = rng.choice(3, size=n, p=[0.40, 0.25, 0.35])
    p = np.empty((n, 3), dtype=np.float32)
    g = kind == 0
    m = int(g.sum())
    rr = 6.0 * np.sqrt(rng.random(m)); th = rng.random(m) * 2 * np.pi
    p[g] = np.stack([rr * np.cos(th), rng.standard_normal(m) * 0.03 - 0.8, rr * np.sin(th)], 1)
    t = kind == 1
    m = int(t.sum())
    p[t] = _blob_mixture(rng, m, 24, (1.6, 1.2, 1.6), 0.18)
    h = kind == 2
    m = int(h.sum())
    rr = rng.uniform(5.0, 9.0, m); th = rng.random(m) * 2 * np.pi
    p[h] = np.stack([rr * np.cos(th), rng.uniform(-0.8, 1.6, m), rr * np.sin(th)], 1)
    rows[:, 0:3] = p
    ls = rng.standard_normal((n, 1), dtype=np.float32) * 1.0 - 4.3
    ls[h] += 0.8
    rows[:, _COL["scale_0"]:_COL["scale_2"] + 1] = ls + rng.standard_normal((n, 3), dtype=np.float32) * 0.5
    rows[:, _COL["opacity"]] = _bimodal_opacity(rng, n)
    return rows


def scene_large(n: int = 50_000_000, seed: int = 5005, background: float = 0.70, chunk: int = 1_000_000,
                threads: int | None = None) -> np.ndarray:
    """C5: the C2 recipe scaled up, 70 % background.  Generated in chunks of `chunk` rows, every chunk from its own
    generator (SeedSequence(seed).spawn) by a pool of threads: the rows depend on (n, seed, background, chunk) only, not
    on the number of threads - 50 M rows (12.4 GB) take seconds instead of minutes of a single generator stream."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    master = np.random.default_rng(seed)
    centers = (master.random((64, 3), dtype=np.float32) - 0.5) * np.asarray((4.0, 2.0, 4.0), dtype=np.float32)
    axes = np.exp(master.standard_normal((64, 3), dtype=np.float32) * 0.5) * np.float32(0.35)
    rows = np.empty((n, ROW_FLOATS), dtype=np.float32)
    nchunks = (n + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(nchunks)
    s0, s2, op = _COL["scale_0"], _COL["scale_2"] + 1, _COL["opacity"]

    def fill(i):
        out = rows[i * chunk:min((i + 1) * chunk, n)]
        m = out.shape[0]
        rng = np.random.default_rng(seeds[i])
        _fill_common(out, rng)
        is_bg = rng.random(m) < background                       # interleaved: ids carry no spatial order
        nbg = int(is_bg.sum())
        nfg = m - nbg
        which = rng.integers(0, 64, size=nfg)
        out[~is_bg, 0:3] = centers[which] + rng.standard_normal((nfg, 3), dtype=np.float32) * axes[which]
        ls = rng.standard_normal((nfg, 1), dtype=np.float32) * 1.1 - 4.6
        out[~is_bg, s0:s2] = ls + rng.standard_normal((nfg, 3), dtype=np.float32) * 0.5
        rad = np.exp(rng.uniform(np.log(4.0), np.log(40.0), nbg)).astype(np.float32)
        d = rng.standard_normal((nbg, 3), dtype=np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        out[is_bg, 0:3] = d * rad[:, None]
        ls = rng.standard_normal((nbg, 1), dtype=np.float32) * 1.0 - 2.5 + np.log(rad / 40.0)[:, None]
        out[is_bg, s0:s2] = ls + rng.standard_normal((nbg, 3), dtype=np.float32) * 0.5
        out[:, op] = _bimodal_opacity(rng, m)

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, range(nchunks)))
    return rows


def write_ply(path: str, rows: np.ndarray, props=PLY_PROPS) -> None:
    rows = np.ascontiguousarray(rows, dtype="<f4")
    assert rows.ndim == 2 and rows.shape[1] == len(props)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {rows.shape[0]}\n"
    header += "".join(f"property float {p}\n" for p in props) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())
