#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own sources (oracle/_ref, built by oracle/build_ref.py from
/root/reference): camera.cc for the matrices, the GLSL shaders executed through glm for every stage, and
cpu_benchmark.cc's std::stable_sort for the order.  Run in the build container (needs /root/reference):

    python oracle/build_ref.py && python tests/golden/make_golden.py

The fixtures are what the `-m "not gpu"` suite pins the oracle against and what the `-m gpu` suite compares the
CUDA path with on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402
from vkgs_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def one(name, rows, w, h, cam, model):
    proj, view, eye = cam
    off = synth.STANDARD_OFFSETS
    pos, cov, op, sh = R.parse_ply(rows, off)
    key, idx = R.rank(pos, proj, view, model)
    skey, sidx = R.sort_key_value(key, idx)
    inv = R.inverse_index(rows.shape[0], sidx)
    inst, indirect = R.projection(pos, cov, op, sh, inv, len(sidx), proj, view, eye, w, h, model)
    image = R.draw(inst, w, h)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), rows=rows, offsets=off, proj=proj, view=view, eye=eye, model=model,
        width=w, height=h, pos=pos, cov=cov, opacity=op, sh=sh, rank_key=key, rank_index=idx, sorted_key=skey,
        sorted_index=sidx, inverse=inv, instances=inst, indirect=indirect, image_f32=image)
    print(name, "N", rows.shape[0], "V", len(sidx), "image mean", image.mean(axis=(0, 1)))


def main():
    assert R.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    eye4 = np.eye(4, dtype=np.float32)
    # A: reference default camera, ball scene
    one("ball_default", synth.scene_c1(n=1500, seed=11), 96, 64, R.camera_default(96, 64), eye4)
    # B: rotated / zoomed / fov-changed / translated camera (Camera::Rotate, Zoom, SetFov, Translate), non-identity model
    model = np.eye(4, dtype=np.float32)
    c, s = np.cos(0.3), np.sin(0.3)
    model[0][0], model[0][2], model[2][0], model[2][2] = c, -s, s, c   # rotation about +Y (column-major m[c][r])
    model *= np.float32(1.25); model[3][3] = 1.0                         # uniform scale
    model[3][0], model[3][1], model[3][2] = 0.1, -0.05, 0.2             # translation
    one("ball_moved", synth.scene_c1(n=1200, seed=12), 80, 60,
        R.camera_ops(80, 60, rot_x=35.0, rot_y=-20.0, zoom=-25.0, fov=np.float32(np.radians(75.0)), tx=30.0, ty=-10.0),
        model.astype(np.float32))
    # C: slab + shell scene (the C2 recipe, tiny), wide image
    one("bicycle_tiny", synth.scene_bicycle(n=2500, seed=13), 128, 72, R.camera_ops(128, 72, zoom=-70.0), eye4)


if __name__ == "__main__":
    main()
