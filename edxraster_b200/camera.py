"""Camera / matrix helpers for the harness (tests, bench, examples).

The reference takes its matrices from EDXUtil's Camera (RealtimeViewer/Main.cpp:39,69-71), which is
absent. These helpers build the three matrices Renderer::SetTransform consumes
(Core/Renderer.cpp:85-92): left-handed view, D3D-style projection (z in [0, w]) and a y-down raster
matrix. Matrices are row-major float32, column-vector convention (clip = P * MV * p).
"""
import math

import numpy as np


def look_at_lh(eye, target, up=(0.0, 1.0, 0.0)):
    eye = np.asarray(eye, np.float64)
    target = np.asarray(target, np.float64)
    up = np.asarray(up, np.float64)
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = x, y, z
    m[0, 3], m[1, 3], m[2, 3] = -x.dot(eye), -y.dot(eye), -z.dot(eye)
    return m.astype(np.float32)


def perspective_lh(fov_y_deg, aspect, near, far):
    ys = 1.0 / math.tan(math.radians(fov_y_deg) * 0.5)
    xs = ys / aspect
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = xs, ys
    m[2, 2] = far / (far - near)
    m[2, 3] = -near * far / (far - near)
    m[3, 2] = 1.0
    return m.astype(np.float32)


def raster_matrix(width, height):
    """NDC [-1,1]^2 (y up) -> raster pixels (y down)."""
    m = np.eye(4)
    m[0, 0], m[0, 3] = width * 0.5, width * 0.5
    m[1, 1], m[1, 3] = -height * 0.5, height * 0.5
    return m.astype(np.float32)


def identity():
    return np.eye(4, dtype=np.float32)


class Camera:
    """Minimal stand-in for EDXUtil's Camera as the viewer uses it (Main.cpp:39)."""

    def __init__(self, eye, target, up, width, height, fov=65.0, near=0.01, far=100.0):
        self.width, self.height = int(width), int(height)
        self.view = look_at_lh(eye, target, up)
        self.proj = perspective_lh(fov, width / float(height), near, far)
        self.raster = raster_matrix(width, height)

    def matrices(self):
        return self.view, self.proj, self.raster
