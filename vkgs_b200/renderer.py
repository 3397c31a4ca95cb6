"""Python host side above the C ABI: the renderer surface of vkgs::Engine (include/vkgs/engine/engine.h:11-26 of the
reference) plus the programmatic camera / viewport / read-back the north-star adds.  numpy only; torch is optional
(device-pointer outputs and streams are passed as integers)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .synth import STANDARD_OFFSETS


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def camera_block(proj, view, eye, model=None) -> L.CameraBlock:
    cb = L.CameraBlock()
    cb.projection[:] = np.ascontiguousarray(proj, np.float32).reshape(16).tolist()
    cb.view[:] = np.ascontiguousarray(view, np.float32).reshape(16).tolist()
    cb.camera_position[:] = np.ascontiguousarray(eye, np.float32).reshape(3).tolist()
    m = np.eye(4, dtype=np.float32) if model is None else np.ascontiguousarray(model, np.float32)
    cb.model[:] = m.reshape(16).tolist()
    return cb


def orbit_camera_block(width, height, fovy=np.radians(60.0), r=2.0, phi=np.radians(45.0), theta=np.radians(45.0),
                       center=(0.0, 0.0, 0.0)) -> L.CameraBlock:
    """vkgs::Camera through the library's own host code (camera.cc): the exact matrices the C++ facade uses."""
    cb = L.CameraBlock()
    c = (C.c_float * 3)(*[float(x) for x in center])
    L.check(L.lib().vkgsb_camera_orbit(int(width), int(height), float(fovy), float(r), float(phi), float(theta), c,
                                       C.byref(cb)))
    return cb


class Renderer:
    def __init__(self, device: int = 0, max_splats: int = 1 << 23, max_width: int = 3840, max_height: int = 2160,
                 max_pairs: int = 0):
        cfg = L.Config(C.sizeof(L.Config), device, max_splats, max_width, max_height, max_pairs)
        h = C.c_void_p()
        L.check(L.lib().vkgsb_create_ex(C.byref(cfg), C.byref(h)))
        self._h = h
        self.width = self.height = 0

    def close(self):
        if getattr(self, "_h", None):
            L.lib().vkgsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- load (Engine::LoadSplats / LoadSplatsAsync, SplatLoadThread::GetProgress / Cancel)
    def load_ply(self, path: str):
        L.check(L.lib().vkgsb_load_ply(self._h, path.encode()))

    def load_ply_async(self, path: str):
        L.check(L.lib().vkgsb_load_ply_async(self._h, path.encode()))

    def load_progress(self):
        t, l, s = C.c_uint32(), C.c_uint32(), C.c_int()
        L.check(L.lib().vkgsb_load_progress(self._h, C.byref(t), C.byref(l), C.byref(s)))
        return dict(total=t.value, loaded=l.value, state=s.value)

    def cancel_load(self):
        L.check(L.lib().vkgsb_cancel_load(self._h))

    def wait_load(self):
        L.check(L.lib().vkgsb_wait_load(self._h))

    def upload_splats(self, rows: np.ndarray, offsets: np.ndarray = STANDARD_OFFSETS):
        rows = np.ascontiguousarray(rows, np.float32)
        offsets = np.ascontiguousarray(offsets, np.uint32)
        if rows.ndim != 2 or offsets.shape != (60,) or rows.shape[1] != int(offsets[59]):
            raise ValueError("rows must be [n, offsets[59]] float32 and offsets a 60-entry table")
        L.check(L.lib().vkgsb_upload_splats(self._h, rows.shape[0], _ptr(rows), _ptr(offsets)))

    # ---- per-frame state
    def set_option(self, option: int, value: int):
        L.check(L.lib().vkgsb_set_option(self._h, option, int(value)))

    def set_blend_mode(self, mode: int):
        self.set_option(L.OPT_BLEND_MODE, mode)

    def set_band(self, y0: int, y1: int):
        self.set_option(L.OPT_BAND_Y0, y0)
        self.set_option(L.OPT_BAND_Y1, y1)

    def set_camera(self, proj=None, view=None, eye=None, model=None, block: L.CameraBlock | None = None):
        cb = block if block is not None else camera_block(proj, view, eye, model)
        L.check(L.lib().vkgsb_set_camera(self._h, C.byref(cb)))

    def set_lines(self, positions=None, colors=None, model=None):
        """Opaque line layer under the splats (the reference's axis / grid, engine.cc:1440-1469): positions [n, 2, 3],
        colors [n, 2, 4] straight alpha, model 4x4 column-major (None = identity).  None / empty removes it."""
        if positions is None or len(positions) == 0:
            L.check(L.lib().vkgsb_set_lines(self._h, 0, None, None, None))
            return
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 6)
        col = np.ascontiguousarray(colors, np.float32).reshape(-1, 8)
        assert pos.shape[0] == col.shape[0]
        m = None if model is None else np.ascontiguousarray(model, np.float32).reshape(16)
        L.check(L.lib().vkgsb_set_lines(self._h, pos.shape[0], _ptr(pos), _ptr(col), _ptr(m) if m is not None else None))

    def set_viewport(self, width: int, height: int):
        L.check(L.lib().vkgsb_set_viewport(self._h, int(width), int(height)))
        self.width, self.height = int(width), int(height)

    # ---- draw
    def draw(self, out: np.ndarray | None = None, stream: int = 0) -> np.ndarray:
        """One frame into host memory (device->host copy included); returns [H, W, 4] uint8."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self.width * self.height * 4
        L.check(L.lib().vkgsb_draw(self._h, _ptr(out), 0, C.c_void_p(stream)))
        return out

    def draw_to_host_ptr(self, host_ptr: int, stream: int = 0):
        L.check(L.lib().vkgsb_draw(self._h, C.c_void_p(host_ptr), 0, C.c_void_p(stream)))

    def draw_device(self, dst_ptr: int = 0, stream: int = 0):
        """One frame left on the device, asynchronous on `stream` (0 = the renderer's stream)."""
        L.check(L.lib().vkgsb_draw(self._h, C.c_void_p(dst_ptr) if dst_ptr else None, 1, C.c_void_p(stream)))

    def draw_batch(self, cameras, out=None, dst_ptr: int = 0, stream: int = 0, stride: int = 0):
        """n views, consecutive frames side by side on the device.  dst_ptr: device destination of view 0, view i goes to
        dst_ptr + i * stride (default stride: one image); asynchronous on `stream`.  Otherwise -> host array [n, H, W, 4]."""
        n = len(cameras)
        arr = (L.CameraBlock * n)(*cameras)
        stride = stride or self.width * self.height * 4
        if dst_ptr:
            L.check(L.lib().vkgsb_draw_batch(self._h, n, arr, C.c_void_p(dst_ptr), stride, 1, C.c_void_p(stream)))
            return None
        if out is None:
            out = np.empty((n, self.height, self.width, 4), np.uint8)
        L.check(L.lib().vkgsb_draw_batch(self._h, n, arr, _ptr(out), stride, 0, C.c_void_p(stream)))
        return out

    def draw_batch_to_host_ptr(self, cameras, host_ptr: int, stream: int = 0):
        """n views into host memory at host_ptr (n * width * height * 4 bytes, ideally pinned): frame i leaves over PCIe
        while frame i + 1 renders; returns when every image has arrived."""
        n = len(cameras)
        arr = (L.CameraBlock * n)(*cameras)
        L.check(L.lib().vkgsb_draw_batch(self._h, n, arr, C.c_void_p(host_ptr), self.width * self.height * 4, 0,
                                         C.c_void_p(stream)))

    def image_device_ptr(self) -> int:
        p = C.c_void_p()
        L.check(L.lib().vkgsb_image_device_ptr(self._h, C.byref(p)))
        return p.value

    def sync(self):
        L.check(L.lib().vkgsb_sync(self._h))

    def stats(self) -> dict:
        s = L.Stats()
        L.check(L.lib().vkgsb_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in L.Stats._fields_}

    def row_histogram(self) -> np.ndarray:
        """Splat centres per image row of the last frame (what balanced band edges are computed from)."""
        rows = np.zeros(self.height, np.uint32)
        L.check(L.lib().vkgsb_row_histogram(self._h, _ptr(rows), self.height))
        return rows

    # ---- band group (one member per GPU)
    def group_export(self) -> bytes:
        h = (C.c_uint8 * 64)()
        L.check(L.lib().vkgsb_group_export(self._h, h))
        return bytes(h)

    def group_join(self, rank: int, world: int, handles, edges):
        """handles: `world` 64-byte handles in rank order (group_export of every member); edges: world + 1 rows."""
        blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(bytes(h) for h in handles))
        e = (C.c_uint32 * (world + 1))(*[int(x) for x in edges])
        L.check(L.lib().vkgsb_group_join(self._h, int(rank), int(world), blob, e))

    def group_leave(self):
        L.check(L.lib().vkgsb_group_leave(self._h))

    # ---- parity taps
    def read_sorted(self):
        cnt = C.c_uint32()
        cap = self.stats()["visible_point_count"]
        keys = np.empty(max(cap, 1), np.uint32); ids = np.empty(max(cap, 1), np.uint32)
        L.check(L.lib().vkgsb_read_sorted(self._h, _ptr(keys), _ptr(ids), cap, C.byref(cnt)))
        return keys[:cnt.value].copy(), ids[:cnt.value].copy()

    def read_instances(self) -> np.ndarray:
        cnt = C.c_uint32()
        cap = self.stats()["visible_point_count"]
        inst = np.empty((max(cap, 1), 12), np.float32)
        L.check(L.lib().vkgsb_read_instances(self._h, _ptr(inst), cap, C.byref(cnt)))
        return inst[:cnt.value].copy()

    def read_order(self) -> np.ndarray:
        """order[i] = index in the loaded file / uploaded rows of stored splat i (the scene is stored in Morton order of
        the centres unless OPT_SPATIAL_ORDER is 0)."""
        cnt = C.c_uint32()
        L.check(L.lib().vkgsb_read_order(self._h, None, 0, C.byref(cnt)))
        order = np.empty(max(cnt.value, 1), np.uint32)
        L.check(L.lib().vkgsb_read_order(self._h, _ptr(order), cnt.value, C.byref(cnt)))
        return order[:cnt.value].copy()

    def read_scene(self):
        """The activated scene in the reference layout, in STORED order (what read_sorted's ids index)."""
        cnt = C.c_uint32()
        L.check(L.lib().vkgsb_read_scene(self._h, None, None, None, None, 0xFFFFFFFF, C.byref(cnt)))
        n = cnt.value
        pos = np.empty((n, 3), np.float32); cov = np.empty((n, 6), np.float32)
        op = np.empty(n, np.float32); sh = np.empty((n, 48), np.uint16)
        if n:
            L.check(L.lib().vkgsb_read_scene(self._h, _ptr(pos), _ptr(cov), _ptr(op), _ptr(sh), n, C.byref(cnt)))
        return pos, cov, op, sh


def reference_overlay(show_axis: bool = True, show_grid: bool = True):
    """The reference viewer's axis and grid as (positions [n,2,3], colors [n,2,4], model[16]): engine.cc:618-680 geometry,
    drawn with model = diag(10, 10, 10, 1) (engine.cc:1444-1448)."""
    pos, col = [], []
    if show_axis:
        for axis, c in enumerate(((1, 0, 0, 1), (0, 1, 0, 1), (0, 0, 1, 1))):
            e = [0.0, 0.0, 0.0]
            e[axis] = 1.0
            pos.append([[0, 0, 0], e])
            col.append([c, c])
    if show_grid:
        g = (0.5, 0.5, 0.5, 1.0)
        for i in range(-10, 11):
            t = np.float32(i) / np.float32(10)
            pos.append([[-1, 0, t], [1, 0, t]])
            col.append([g, g])
            pos.append([[t, 0, -1], [t, 0, 1]])
            col.append([g, g])
    model = np.diag([10.0, 10.0, 10.0, 1.0]).astype(np.float32).T.reshape(16)
    return (np.asarray(pos, np.float32).reshape(-1, 2, 3), np.asarray(col, np.float32).reshape(-1, 2, 4), model)


def external_alloc(device: int, nbytes: int):
    """Exportable device memory (the stand-in for a VkImage's memory): (handle, fd, device pointer); the caller closes fd."""
    h, fd, p = C.c_void_p(), C.c_int(-1), C.c_void_p()
    L.check(L.lib().vkgsb_external_alloc(int(device), int(nbytes), C.byref(h), C.byref(fd), C.byref(p)))
    return h, fd.value, p.value


def external_import(device: int, fd: int, nbytes: int, handle_type: int = L.EXTERNAL_OPAQUE_FD):
    """Memory another API / process exported as a file descriptor -> (handle, device pointer usable as a frame destination)."""
    h, p = C.c_void_p(), C.c_void_p()
    L.check(L.lib().vkgsb_external_import(int(device), int(fd), int(nbytes), int(handle_type), C.byref(h), C.byref(p)))
    return h, p.value


def external_release(handle):
    L.check(L.lib().vkgsb_external_release(handle))


def group_join_local(members, edges):
    """Band group of renderers living in this process (member g draws rows [edges[g], edges[g + 1]))."""
    n = len(members)
    arr = (C.c_void_p * n)(*[m._h for m in members])
    e = (C.c_uint32 * (n + 1))(*[int(x) for x in edges])
    L.check(L.lib().vkgsb_group_join_local(arr, n, e))


def shared_create(device: int, nbytes: int):
    """A device buffer other processes of the node can map: (device pointer, 64-byte handle)."""
    p = C.c_void_p()
    h = (C.c_uint8 * 64)()
    L.check(L.lib().vkgsb_shared_create(int(device), int(nbytes), C.byref(p), h))
    return p.value, bytes(h)


def shared_open(device: int, handle: bytes) -> int:
    p = C.c_void_p()
    h = (C.c_uint8 * 64).from_buffer_copy(handle)
    L.check(L.lib().vkgsb_shared_open(int(device), h, C.byref(p)))
    return p.value


def shared_close(device: int, ptr: int):
    L.check(L.lib().vkgsb_shared_close(int(device), C.c_void_p(ptr)))


def shared_read(device: int, ptr: int, offset: int, shape) -> np.ndarray:
    out = np.empty(shape, np.uint8)
    L.check(L.lib().vkgsb_shared_read(int(device), C.c_void_p(ptr), int(offset), out.size, _ptr(out)))
    return out


def shared_destroy(device: int, ptr: int):
    L.check(L.lib().vkgsb_shared_destroy(int(device), C.c_void_p(ptr)))


def device_count() -> int:
    c = C.c_int()
    L.check(L.lib().vkgsb_device_count(C.byref(c)))
    return c.value


def sort_storage_bytes(max_n: int) -> int:
    b = C.c_size_t()
    L.check(L.lib().vkgsb_sort_storage_bytes(int(max_n), C.byref(b)))
    return b.value


def sort_key_value_indirect(stream: int, max_n: int, d_count: int, d_keys: int, d_values: int, d_storage: int):
    """Device pointers as integers (e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
    L.check(L.lib().vkgsb_sort_key_value_indirect(C.c_void_p(stream), int(max_n), C.c_void_p(d_count),
                                                  C.c_void_p(d_keys), C.c_void_p(d_values), C.c_void_p(d_storage)))
