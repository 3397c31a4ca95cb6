#!/usr/bin/env python
"""CPU model (numpy, float32) of the load-time spatial order and the tile-box classification of the cull
(vkgs_b200/csrc/spatial.cu, k_cull_classify in project.cu): how many tiles of 256 splats are skipped / taken whole /
tested per splat on a configuration's orbit, and that the classification is conservative against the per-splat test.

  python tools/spatial_model.py [--config c2] [--views 0,21,42] [--n N]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
F = np.float32
EPS = F(2e-5)


def spatial_order(pos, lmax=None, bits=10):
    """Stored order: stable sort by (size class, 3 x `bits` Morton code of the centres quantised over mean +- 3 sigma,
    clamped to the bounding box); size class from lmax = the largest eigenvalue of the 3-D covariance: 3 sigma above 1/16
    or 1/128 of the window's extent go first.  float64 here; the device does the same with integer moments
    (deterministic)."""
    lo, hi = pos.min(0).astype(np.float64), pos.max(0).astype(np.float64)
    ext = np.maximum(hi - lo, 1e-30)
    u = (pos.astype(np.float64) - lo) / ext
    m, s = u.mean(0), u.std(0)
    a, b = np.maximum(m - 3 * s, 0.0), np.minimum(m + 3 * s, 1.0)
    q = np.clip(((u - a) / np.maximum(b - a, 1e-30) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(pos), np.int64)
    for bit in range(bits):
        for ax in range(3):
            key |= ((q[:, ax] >> bit) & 1) << (3 * bit + ax)
    if lmax is not None:
        w = ((b - a) * ext).max()
        r2 = 9.0 * np.asarray(lmax, np.float64)
        cls = (r2 * 16384.0 > w * w).astype(np.int64) + (r2 * 256.0 > w * w).astype(np.int64)
        key |= (2 - cls) << (3 * bits)
    return np.argsort(key, kind="stable")


def tile_boxes(pos, tile=256):
    n = len(pos)
    nt = (n + tile - 1) // tile
    pad = nt * tile - n
    p = np.concatenate([pos, np.repeat(pos[-1:], pad, 0)]) if pad else pos
    p = p.reshape(nt, tile, 3)
    return p.min(1), p.max(1)


def classify(pvm, lo, hi):
    """0 = no splat of the tile is visible, 1 = every splat is, 2 = test per splat.  pvm: [4][4] column-major (m[c][r])."""
    M = np.asarray(pvm, F).reshape(4, 4)
    amax = np.maximum(np.abs(lo), np.abs(hi))

    def rng(k):  # k[4] coefficients of a linear functional over the box
        mn = k[3] + sum(np.where(k[j] >= 0, k[j] * lo[:, j], k[j] * hi[:, j]) for j in range(3))
        mx = k[3] + sum(np.where(k[j] >= 0, k[j] * hi[:, j], k[j] * lo[:, j]) for j in range(3))
        return mn.astype(F), mx.astype(F)

    row = [M[:, i] for i in range(4)]
    mag = [np.abs(row[i][3]) + sum(np.abs(row[i][j]) * amax[:, j] for j in range(3)) for i in range(4)]
    c3mn, c3mx = rng(row[3])
    fs = [(row[3] - row[0], mag[3] + mag[0]), (row[3] + row[0], mag[3] + mag[0]), (row[3] - row[1], mag[3] + mag[1]),
          (row[3] + row[1], mag[3] + mag[1]), (row[2], mag[2]), (row[3] - row[2], mag[3] + mag[2])]
    inside = np.ones(len(lo), bool)
    out_front = np.zeros(len(lo), bool)
    out_back = np.zeros(len(lo), bool)
    for k, mg in fs:
        mn, mx = rng(k)
        m = EPS * mg
        inside &= mn >= m
        out_front |= mx < -m
        out_back |= mn > m
    m3 = EPS * mag[3]
    front, back = c3mn > m3, c3mx < -m3
    outside = np.where(front, out_front, np.where(back, out_back, out_front & out_back))
    cls = np.full(len(lo), 2, np.int32)
    cls[outside] = 0
    cls[front & inside & ~outside] = 1
    return cls


def band_params(P, V, width, height, model=None):
    """bc_a, bc_b, bc_p of fill_params (renderer.cu) for a uniformly scaled rotation as the model matrix."""
    Vm = np.asarray(V, np.float64).reshape(4, 4).T[:3, :3]
    Mm = np.eye(3) if model is None else np.asarray(model, np.float64).reshape(4, 4).T[:3, :3]
    W = Vm @ Mm
    G = W.T @ W
    w2 = min((W * W).sum(), np.abs(G).sum(1).max()) * (1.0 + 1e-6)
    hh = 0.5 * height
    Pm = np.asarray(P, np.float64).reshape(4, 4)
    return F(9.0 * hh * hh * w2), F(9.0 * hh * hh * (1.0 / width ** 2 + 1.0 / height ** 2)), F(Pm[0, 0] ** 2 + Pm[1, 1] ** 2)


def band_miss_splats(pvm, pos, lmax, bc, height, y0, y1):
    """band_miss_rows (project.cu) per splat, float32."""
    M = np.asarray(pvm, F).reshape(4, 4)
    c = (pos @ M[:3, :] + M[3, :]).astype(F)
    with np.errstate(all="ignore"):
        iw = (F(1) / c[:, 3]).astype(F)
        xn, yn = c[:, 0] * iw, c[:, 1] * iw
        hh = F(0.5 * height)
        cpy = yn * hh + (hh - F(0.5))
        d = np.maximum(np.maximum(F(y0) - cpy, cpy - (F(y1) - F(1))), F(0)) - F(2)
        pj2 = (bc[2] + xn * xn + yn * yn) * (iw * iw)
        bound = (bc[0] * lmax * pj2 + bc[1]) * F(1.01)
        return (d > 0) & (d * d > bound)


def band_classify(pvm, lo, hi, trmax, bc, height, y0, y1):
    """The band part of classify_tile (project.cu): True where no footprint of the tile can reach rows [y0, y1)."""
    M = np.asarray(pvm, F).reshape(4, 4)
    amax = np.maximum(np.abs(lo), np.abs(hi))

    def rng(i):
        k = M[:, i]
        mn = k[3] + sum(np.where(k[j] >= 0, k[j] * lo[:, j], k[j] * hi[:, j]) for j in range(3))
        mx = k[3] + sum(np.where(k[j] >= 0, k[j] * hi[:, j], k[j] * lo[:, j]) for j in range(3))
        return mn.astype(F), mx.astype(F)

    (x0, x1), (yy0, yy1), (w0, w1) = rng(0), rng(1), rng(3)
    mag3 = np.abs(M[3, 3]) + sum(np.abs(M[j, 3]) * amax[:, j] for j in range(3))
    front = w0 > EPS * mag3
    with np.errstate(all="ignore"):
        iwmax, iwmin = F(1) / w0, F(1) / w1
        xhi = np.where(x1 >= 0, x1 * iwmax, x1 * iwmin); xlo = np.where(x0 >= 0, x0 * iwmin, x0 * iwmax)
        yhi = np.where(yy1 >= 0, yy1 * iwmax, yy1 * iwmin); ylo = np.where(yy0 >= 0, yy0 * iwmin, yy0 * iwmax)
        hh = F(0.5 * height)
        cpy_hi, cpy_lo = yhi * hh + (hh - F(0.5)), ylo * hh + (hh - F(0.5))
        pj2 = (bc[2] + np.maximum(xlo * xlo, xhi * xhi) + np.maximum(ylo * ylo, yhi * yhi)) * (iwmax * iwmax)
        bound = (bc[0] * trmax * pj2 + bc[1]) * F(1.03)
        d = np.maximum(np.maximum(F(y0) - cpy_hi, cpy_lo - (F(y1) - F(1))), F(0)) - F(2.1)
        return front & (d > 0) & (d * d > bound)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--views", default="0,21,42")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--no-order", action="store_true")
    a = ap.parse_args()
    import bench
    from oracle import oracle as O
    cfg = dict(bench.CONFIGS[a.config])
    if a.n:
        cfg["n_splats"] = a.n
    rows = bench.make_scene(cfg)
    pos = np.ascontiguousarray(rows[:, 0:3])
    del rows
    if not a.no_order:
        pos = pos[spatial_order(pos)]
    lo, hi = tile_boxes(pos)
    for v in [int(x) for x in a.views.split(",")]:
        P, V, E = bench.view_camera(cfg, v)
        pvm = O.compose_pvm(P, V)
        cls = classify(pvm, lo, hi)
        M = np.asarray(pvm, F).reshape(4, 4)
        c = (pos @ M[:3, :] + M[3, :]).astype(F)
        with np.errstate(all="ignore"):
            ndc = c[:, :3] / c[:, 3:4]
        vis = (np.abs(ndc[:, 0]) <= 1) & (np.abs(ndc[:, 1]) <= 1) & (ndc[:, 2] >= 0) & (ndc[:, 2] <= 1)
        nt = len(lo)
        pad = nt * 256 - len(pos)
        vt = np.concatenate([vis, np.zeros(pad, bool)]).reshape(nt, 256).sum(1)
        full = np.full(nt, 256)
        full[-1] -= pad
        bad_out = int(((cls == 0) & (vt != 0)).sum())
        bad_in = int(((cls == 1) & (vt != full)).sum())
        print(f"view {v}: visible {int(vis.sum())} of {len(pos)}; tiles {nt}: outside {int((cls == 0).sum())}, inside "
              f"{int((cls == 1).sum())}, per-splat {int((cls == 2).sum())} ({100.0 * (cls == 2).mean():.1f} %); "
              f"non-conservative: {bad_out} outside, {bad_in} inside; runs of visible ids: "
              f"{int((np.diff(vis.astype(np.int8)) == 1).sum())}")


if __name__ == "__main__":
    main()
