"""pygs: drop-in for the reference's Python package (binding/python/pygs/__init__.py:14-32): show / load / close with
the same behaviour (load() raises FileNotFoundError for a missing file), running on libvkgsb's CUDA renderer instead of
the Vulkan viewer.  Headless extensions: render(), set_orbit(), wait_loaded(), stats().

The native module pygs/_pygs_cpp*.so is built by `python -m vkgs_b200.build`; there is no pure-Python fallback.
"""
import errno
import os

import numpy as np

try:
    from . import _pygs_cpp as _C
except ImportError as e:  # fail loudly: the binding is native code
    raise ImportError("pygs._pygs_cpp is not built: run `python -m vkgs_b200.build`") from e


def show():
    _C.show()


def load(ply_filepath):
    ply_filepath = os.path.abspath(ply_filepath)
    if not os.path.exists(ply_filepath):
        raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), ply_filepath)
    print(f"load {ply_filepath}")
    _C.ensure_engine()   # headless use: load() works without a prior show()
    _C.load(ply_filepath)


def close():
    _C.close()


def set_orbit(center=(0.0, 0.0, 0.0), r=2.0, phi=np.radians(45.0), theta=np.radians(45.0), fovy=0.0):
    """Orbit pose of vkgs::Camera (camera.h:42-52 defaults); fovy = 0 keeps the current field of view."""
    _C.set_orbit(float(center[0]), float(center[1]), float(center[2]), float(r), float(phi), float(theta), float(fovy))


def wait_loaded():
    _C.wait_loaded()


def render(width=1600, height=900):
    """One offscreen frame at the current camera -> [H, W, 4] uint8 (RGBA)."""
    return np.frombuffer(_C.render(int(width), int(height)), np.uint8).reshape(int(height), int(width), 4)


def stats():
    return _C.stats()


__all__ = ["show", "load", "close", "set_orbit", "wait_loaded", "render", "stats"]
