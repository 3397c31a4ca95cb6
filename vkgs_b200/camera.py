"""Orbit camera: host-side mirror of vkgs::Camera (include/vkgs/scene/camera.h:8-58,
src/vkgs/scene/camera.cc:25-70) producing the per-frame parameter block of
src/vkgs/vulkan/shader/uniforms.h:10-15 (column-major float32 matrices, like glm).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _normalize(v):
    v = np.asarray(v, dtype=F)
    return v * F(1.0 / math.sqrt(float(np.dot(v, v))))


def perspective_rh_no(fovy, aspect, near, far) -> np.ndarray:
    """glm::perspective (RH, depth -1..1): third_party/glm/glm/ext/matrix_clip_space.inl:249-262."""
    t = F(math.tan(float(F(fovy)) / 2.0))
    m = np.zeros((4, 4), dtype=F)  # m[c][r]
    m[0][0] = F(1) / (F(aspect) * t)
    m[1][1] = F(1) / t
    m[2][2] = -(F(far) + F(near)) / (F(far) - F(near))
    m[2][3] = -F(1)
    m[3][2] = -(F(2) * F(far) * F(near)) / (F(far) - F(near))
    return m


def look_at_rh(eye, center, up) -> np.ndarray:
    """glm::lookAt (RH): third_party/glm/glm/ext/matrix_transform.inl lookAtRH."""
    eye = np.asarray(eye, dtype=F); center = np.asarray(center, dtype=F)
    f = _normalize(center - eye)
    s = _normalize(np.cross(f, np.asarray(up, dtype=F)).astype(F))
    u = np.cross(s, f).astype(F)
    m = np.eye(4, dtype=F)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -np.dot(s, eye)
    m[3][1] = -np.dot(u, eye)
    m[3][2] = np.dot(f, eye)
    return m


class Camera:
    """Same state, defaults and mouse operations as vkgs::Camera (camera.h:41-57)."""

    MIN_FOV = math.radians(40.0)
    MAX_FOV = math.radians(100.0)

    def __init__(self, width=256, height=256):
        self.width, self.height = int(width), int(height)
        self.fovy = F(math.radians(60.0))
        self.near, self.far = F(0.01), F(100.0)
        self.center = np.zeros(3, dtype=F)
        self.r = F(2.0)
        self.phi = F(math.radians(45.0))
        self.theta = F(math.radians(45.0))
        self.rotation_sensitivity = F(0.01)
        self.translation_sensitivity = F(0.002)
        self.zoom_sensitivity = F(0.01)
        self.dolly_zoom_sensitivity = F(math.radians(1.0))

    def set_window_size(self, width, height):
        self.width, self.height = int(width), int(height)

    def set_fov(self, fov):  # camera.cc:18-23, dolly zoom
        self.r = F(self.r * F(math.tan(self.fovy / 2.0)) / F(math.tan(fov / 2.0)))
        self.fovy = F(fov)

    def projection_matrix(self) -> np.ndarray:  # camera.cc:25-35
        aspect = F(self.width) / F(self.height)
        p = perspective_rh_no(self.fovy, aspect, self.near, self.far)
        conv = np.eye(4, dtype=F)
        conv[1][1] = -1.0
        conv[2][2] = 0.5
        conv[3][2] = 0.5
        # column-major product conv * p: as row-major numpy arrays holding m[c][r], (A*B)[c] = sum_k A[k]*B[c][k]
        return (p @ conv).astype(F)

    def eye(self) -> np.ndarray:  # camera.cc:39-45
        sp, cp = F(math.sin(self.phi)), F(math.cos(self.phi))
        st, ct = F(math.sin(self.theta)), F(math.cos(self.theta))
        return (self.center + self.r * np.array([sp * st, cp, sp * ct], dtype=F)).astype(F)

    def view_matrix(self) -> np.ndarray:  # camera.cc:37
        return look_at_rh(self.eye(), self.center, (0.0, 1.0, 0.0))

    def rotate(self, x, y):  # camera.cc:47-51
        self.theta = F(self.theta - self.rotation_sensitivity * F(x))
        eps = math.radians(0.1)
        self.phi = F(min(max(float(self.phi - self.rotation_sensitivity * F(y)), eps), math.pi - eps))

    def translate(self, x, y, z=0.0):  # camera.cc:53-63
        sp, cp = math.sin(self.phi), math.cos(self.phi)
        st, ct = math.sin(self.theta), math.cos(self.theta)
        d = (-x * np.array([ct, 0.0, -st]) + y * np.array([-cp * st, sp, -cp * ct])
             - z * np.array([sp * st, cp, sp * ct]))
        self.center = (self.center + F(self.translation_sensitivity * self.r) * d.astype(F)).astype(F)

    def zoom(self, x):  # camera.cc:65
        self.r = F(self.r / F(math.exp(self.zoom_sensitivity * x)))

    def dolly_zoom(self, scroll):  # camera.cc:67-70
        self.set_fov(min(max(float(self.fovy - scroll * self.dolly_zoom_sensitivity), self.MIN_FOV), self.MAX_FOV))


def orbit(width, height, r=2.0, phi_deg=45.0, theta_deg=45.0, center=(0, 0, 0), fovy_deg=60.0) -> Camera:
    cam = Camera(width, height)
    cam.r = F(r); cam.phi = F(math.radians(phi_deg)); cam.theta = F(math.radians(theta_deg))
    cam.center = np.asarray(center, dtype=F); cam.fovy = F(math.radians(fovy_deg))
    return cam
