"""ctypes front-end of the CPU oracle (oracle/vkgs_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvkgs_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vkgs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class Camera(C.Structure):
    _fields_ = [
        ("proj", C.c_float * 16),
        ("view", C.c_float * 16),
        ("model", C.c_float * 16),
        ("eye", C.c_float * 3),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
    ]


class FrameStats(C.Structure):
    _fields_ = [("visible", C.c_uint32), ("ms_cull", C.c_double), ("ms_sort", C.c_double),
                ("ms_project", C.c_double), ("ms_raster", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.vko_cull.restype = C.c_uint32
        _lib.vko_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_camera(proj, view, eye, width, height, model=None) -> Camera:
    cam = Camera()
    cam.proj[:] = _f32(proj).reshape(16).tolist()
    cam.view[:] = _f32(view).reshape(16).tolist()
    cam.model[:] = (_f32(model).reshape(16) if model is not None else np.eye(4, dtype=np.float32).reshape(16)).tolist()
    cam.eye[:] = _f32(eye).reshape(3).tolist()
    cam.width, cam.height = int(width), int(height)
    return cam


class Scene:
    """Activated scene in the reference's SoA layout (engine.cc:1639-1651)."""

    def __init__(self, pos, cov, opacity, sh):
        self.pos = _f32(pos).reshape(-1, 3)
        self.cov = _f32(cov).reshape(-1, 6)
        self.opacity = _f32(opacity).reshape(-1)
        self.sh = np.ascontiguousarray(sh, dtype=np.uint16).reshape(-1, 48)
        self.n = self.pos.shape[0]


def activate(rows: np.ndarray, offsets: np.ndarray) -> Scene:
    rows = _f32(rows)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    n = rows.shape[0]
    assert rows.shape[1] == int(offsets[59])
    pos = np.empty((n, 3), np.float32); cov = np.empty((n, 6), np.float32)
    op = np.empty(n, np.float32); sh = np.empty((n, 48), np.uint16)
    lib().vko_activate(C.c_uint32(n), _p(rows), _p(offsets), _p(pos), _p(cov), _p(op), _p(sh))
    return Scene(pos, cov, op, sh)


def compose_pvm(proj, view, model=None) -> np.ndarray:
    out = np.empty(16, np.float32)
    m = _f32(model).reshape(16) if model is not None else np.eye(4, dtype=np.float32).reshape(16)
    lib().vko_compose_pvm(_p(_f32(proj).reshape(16)), _p(_f32(view).reshape(16)), _p(m), _p(out))
    return out


def cull(scene: Scene, pvm: np.ndarray):
    keys = np.empty(scene.n, np.uint32); ids = np.empty(scene.n, np.uint32)
    v = lib().vko_cull(C.c_uint32(scene.n), _p(scene.pos), _p(_f32(pvm).reshape(16)), _p(keys), _p(ids))
    return keys[:v].copy(), ids[:v].copy()


def sort_pairs(keys: np.ndarray, vals: np.ndarray):
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy(); v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    lib().vko_sort_pairs(C.c_uint32(k.shape[0]), _p(k), _p(v))
    return k, v


def inverse_index(n: int, index: np.ndarray) -> np.ndarray:
    index = np.ascontiguousarray(index, dtype=np.uint32)
    inv = np.empty(n, np.int32)
    lib().vko_inverse_index(C.c_uint32(n), C.c_uint32(index.shape[0]), _p(index), _p(inv))
    return inv


def project(scene: Scene, ids: np.ndarray, cam: Camera, variant: int = 0) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    inst = np.empty((ids.shape[0], 12), np.float32)
    lib().vko_project(C.c_uint32(ids.shape[0]), _p(ids), _p(scene.pos), _p(scene.cov), _p(scene.opacity),
                      _p(scene.sh), C.byref(cam), C.c_int(variant), _p(inst))
    return inst


def raster(inst: np.ndarray, width: int, height: int, mode: int = 0, tile: int = 16, want_float: bool = False):
    inst = _f32(inst).reshape(-1, 12)
    out = np.empty((height, width, 4), np.uint8)
    fout = np.empty((height, width, 4), np.float32) if want_float else None
    lib().vko_raster(C.c_uint32(inst.shape[0]), _p(inst), C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile),
                     C.c_int(mode), _p(out), _p(fout) if want_float else None)
    return (out, fout) if want_float else out


def raster_rows(inst: np.ndarray, width: int, height: int, row0: int, row1: int, mode: int = 0, tile: int = 16):
    """Only tile bands overlapping rows [row0,row1) are rasterised (bounded CPU sample for bench.py)."""
    inst = _f32(inst).reshape(-1, 12)
    out = np.zeros((height, width, 4), np.uint8)
    lib().vko_raster_rows(C.c_uint32(inst.shape[0]), _p(inst), C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile),
                          C.c_int(mode), C.c_uint32(row0), C.c_uint32(row1), _p(out), None)
    return out


def raster_band_list(inst: np.ndarray, width: int, height: int, bands, mode: int = 0, tile: int = 16):
    """The listed tile bands (rows [b * tile, b * tile + tile)), one OpenMP thread per band; other rows stay zero."""
    inst = _f32(inst).reshape(-1, 12)
    bands = np.ascontiguousarray(bands, dtype=np.uint32)
    out = np.zeros((height, width, 4), np.uint8)
    lib().vko_raster_band_list(C.c_uint32(inst.shape[0]), _p(inst), C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile),
                               C.c_int(mode), C.c_uint32(bands.shape[0]), _p(bands), _p(out))
    return out


def raster_lines(positions, colors, pvm, width: int, height: int):
    """The opaque line layer (axis / grid of the reference viewer): (depth f32 [H,W], rgba u8 [H,W,4])."""
    pos = _f32(positions).reshape(-1, 6)
    col = _f32(colors).reshape(-1, 8)
    assert pos.shape[0] == col.shape[0]
    pvm = _f32(pvm).reshape(16)
    depth = np.empty((height, width), np.float32)
    rgba = np.empty((height, width, 4), np.uint8)
    lib().vko_raster_lines(C.c_uint32(pos.shape[0]), _p(pos), _p(col), _p(pvm), C.c_uint32(width), C.c_uint32(height),
                           _p(depth), _p(rgba))
    return depth, rgba


def raster_layer(inst: np.ndarray, width: int, height: int, layer_depth, layer_rgba, mode: int = 0, tile: int = 16):
    """Splats over an opaque layer: depth-tested LESS against layer_depth, blended over layer_rgba."""
    inst = _f32(inst).reshape(-1, 12)
    out = np.zeros((height, width, 4), np.uint8)
    ld = _f32(layer_depth).reshape(height, width)
    lc = np.ascontiguousarray(layer_rgba, np.uint8).reshape(height, width, 4)
    lib().vko_raster_rows_layer(C.c_uint32(inst.shape[0]), _p(inst), C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile),
                                C.c_int(mode), C.c_uint32(0), C.c_uint32(height), _p(ld), _p(lc), _p(out), None)
    return out


def render(scene: Scene, cam: Camera, mode: int = 0, tile: int = 16):
    """Whole frame; returns dict(image, keys, ids, inst, stats)."""
    keys = np.empty(scene.n, np.uint32); ids = np.empty(scene.n, np.uint32)
    inst = np.empty((scene.n, 12), np.float32)
    out = np.empty((cam.height, cam.width, 4), np.uint8)
    st = FrameStats()
    lib().vko_render(C.c_uint32(scene.n), _p(scene.pos), _p(scene.cov), _p(scene.opacity), _p(scene.sh), C.byref(cam),
                     C.c_uint32(tile), C.c_int(mode), _p(keys), _p(ids), _p(inst), _p(out), C.byref(st))
    v = st.visible
    return dict(image=out, keys=keys[:v].copy(), ids=ids[:v].copy(), inst=inst[:v].copy(),
                stats=dict(visible=v, ms_cull=st.ms_cull, ms_sort=st.ms_sort, ms_project=st.ms_project,
                           ms_raster=st.ms_raster))


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().vko_set_num_threads(C.c_int(n))
    return num_threads()


def num_threads() -> int:
    return int(lib().vko_num_threads())
