"""ctypes front-end of oracle/_ref (the reference's own sources compiled by oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY.  `available()` is False when oracle/_ref has not been built (e.g. a checkout
without /root/reference); callers skip.  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_libs = {}


def _lib(name):
    if name not in _libs:
        path = os.path.join(_DIR, name)
        _libs[name] = C.CDLL(path) if os.path.exists(path) else None
    return _libs[name]


def available() -> bool:
    return all(_lib(n) is not None for n in ("libref_camera.so", "libref_sort.so", "libref_shaders.so"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def camera_default(w, h):
    p = np.empty(16, np.float32); v = np.empty(16, np.float32); e = np.empty(3, np.float32)
    _lib("libref_camera.so").ref_camera_default(C.c_uint(w), C.c_uint(h), _p(p), _p(v), _p(e))
    return p.reshape(4, 4), v.reshape(4, 4), e


def camera_ops(w, h, rot_x=0.0, rot_y=0.0, zoom=0.0, fov=-1.0, tx=0.0, ty=0.0, tz=0.0):
    p = np.empty(16, np.float32); v = np.empty(16, np.float32); e = np.empty(3, np.float32)
    f = C.c_float
    _lib("libref_camera.so").ref_camera_ops(C.c_uint(w), C.c_uint(h), f(rot_x), f(rot_y), f(zoom), f(fov), f(tx), f(ty),
                                            f(tz), _p(p), _p(v), _p(e))
    return p.reshape(4, 4), v.reshape(4, 4), e


def sort_key_value(keys, vals):
    keys = np.ascontiguousarray(keys, np.uint32); vals = np.ascontiguousarray(vals, np.uint32)
    ok = np.empty_like(keys); ov = np.empty_like(vals)
    _lib("libref_sort.so").ref_sort_key_value(C.c_uint(keys.shape[0]), _p(keys), _p(vals), _p(ok), _p(ov))
    return ok, ov


def parse_ply(rows, offsets):
    rows = _f32(rows); offsets = np.ascontiguousarray(offsets, np.uint32)
    n = rows.shape[0]
    pos = np.empty((n, 3), np.float32); cov = np.empty((n, 6), np.float32)
    op = np.empty(n, np.float32); sh = np.empty((n, 48), np.uint16)
    _lib("libref_shaders.so").ref_parse_ply(C.c_uint(n), _p(offsets), _p(rows), _p(pos), _p(cov), _p(op), _p(sh))
    return pos, cov, op, sh


def rank(pos, proj, view, model):
    pos = _f32(pos); n = pos.shape[0]
    key = np.empty(n, np.uint32); idx = np.empty(n, np.uint32)
    L = _lib("libref_shaders.so"); L.ref_rank.restype = C.c_uint
    v = L.ref_rank(C.c_uint(n), _p(_f32(proj).reshape(16)), _p(_f32(view).reshape(16)), _p(_f32(model).reshape(16)),
                   _p(pos), _p(key), _p(idx))
    return key[:v].copy(), idx[:v].copy()


def inverse_index(n, index):
    index = np.ascontiguousarray(index, np.uint32); inv = np.empty(n, np.int32)
    _lib("libref_shaders.so").ref_inverse_index(C.c_uint(n), C.c_uint(index.shape[0]), _p(index), _p(inv))
    return inv


def projection(pos, cov, opacity, sh, inverse, visible, proj, view, eye, width, height, model):
    pos = _f32(pos); n = pos.shape[0]
    inst = np.full((n, 12), np.nan, np.float32); ind = np.zeros(12, np.uint32)
    _lib("libref_shaders.so").ref_projection(
        C.c_uint(n), C.c_uint(visible), _p(_f32(proj).reshape(16)), _p(_f32(view).reshape(16)), _p(_f32(eye)),
        C.c_uint(width), C.c_uint(height), _p(_f32(model).reshape(16)), _p(pos), _p(_f32(cov)), _p(_f32(opacity)),
        _p(np.ascontiguousarray(sh, np.uint16)), _p(np.ascontiguousarray(inverse, np.int32)), _p(inst), _p(ind))
    return inst[:visible].copy(), ind


def draw(inst, width, height):
    inst = _f32(inst).reshape(-1, 12)
    out = np.empty((height, width, 4), np.float32)
    _lib("libref_shaders.so").ref_draw(_p(inst), C.c_uint(inst.shape[0]), C.c_uint(width), C.c_uint(height), _p(out))
    return out
