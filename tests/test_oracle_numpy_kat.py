"""A second, scalar restatement of the per-pixel arithmetic in numpy float32 (no SIMD, no tiling, no shared
code with the oracle), evaluated for every covered pixel of one unclipped triangle and compared bit for bit
with the oracle's depth and colour. It follows the reference formulas directly:
setup RasterTriangle.h:27-60, barycentrics/depth :324-337, Interpolate Shader.h:142-170,
Blinn-Phong Shader.h:246-282, pack Renderer.cpp:295-301 (with the EDXUtil definitions of DESIGN.md §2)."""
import numpy as np

from edxraster_b200 import camera as cam, scenes
from oracle import orc

f32 = np.float32


def row_dot(m, x, y, z):
    return f32(f32(f32(f32(m[0] * x) + f32(m[1] * y)) + f32(m[2] * z)) + m[3])


def test_scalar_numpy_restatement_matches_the_oracle():
    W, H = 96, 64
    c = cam.Camera((0.3, 0.2, -3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), W, H, 60.0, 0.1, 50.0)
    P = np.array([[-0.9, -0.7, 0.4], [0.1, 1.1, -0.3], [1.2, -0.5, 0.2]], f32)
    N = np.array([[0.2, 0.1, -1.0], [-0.3, 0.4, -0.8], [0.5, -0.2, -0.9]], f32)
    v = np.zeros((3, 8), f32)
    v[:, 0:3], v[:, 3:6] = P, N
    idx = np.array([[0, 1, 2]], np.uint32)
    o = orc.Oracle(W, H, 1)
    o.set_transform(c.view, c.proj, c.raster)
    o.set_shader(scenes.SHADER_BLINN_PHONG)
    o.render(v, idx)
    if (o.winner() != 0xFFFFFFFF).sum() == 0:          # wrong winding for this camera: flip
        idx = np.array([[0, 2, 1]], np.uint32)
        P, N = P[[0, 2, 1]], N[[0, 2, 1]]
        o.render(v, idx)
    mvp, eye, light = o.derived()
    depth, color, winner = o.depth()[::-1], o.color()[::-1], o.winner()[::-1]
    assert (winner != 0xFFFFFFFF).sum() > 300

    R = c.raster.astype(f32)
    clip = np.array([[row_dot(mvp[r], *p) for r in range(4)] for p in P], f32)
    snapped, z, invw = [], [], []
    for cx, cy, cz, cw in clip:
        ax, ay, az = f32(cx / cw), f32(cy / cw), f32(cz / cw)
        x, y, w = row_dot(R[0], ax, ay, az), row_dot(R[1], ax, ay, az), row_dot(R[3], ax, ay, az)
        if w != f32(1.0):
            x, y = f32(x / w), f32(y / w)
        snapped.append((int(np.trunc(np.float64(x) * 16.0)), int(np.trunc(np.float64(y) * 16.0))))
        iw = f32(f32(1.0) / cw)
        invw.append(iw)
        z.append(f32(cz * iw))
    (v0x, v0y), (v1x, v1y), (v2x, v2y) = snapped
    B0, C0, B1, C1, B2, C2 = v0y - v1y, v1x - v0x, v1y - v2y, v2x - v1x, v2y - v0y, v0x - v2x
    det = C2 * B1 - C1 * B2
    assert det > 0
    inv_det = f32(f32(1.0) / f32(det))
    tl = lambda a, b: -1 if (b[1] > a[1] or (a[1] == b[1] and a[0] > b[0])) else 0
    bias = (tl(snapped[0], snapped[1]), tl(snapped[1], snapped[2]), tl(snapped[2], snapped[0]))

    checked = 0
    for py in range(H):
        for px in range(W):
            cx, cy = 16 * px + 8, 16 * py + 8
            e0 = B0 * (cx - v0x) + C0 * (cy - v0y) + bias[0]
            e1 = B1 * (cx - v1x) + C1 * (cy - v1y) + bias[1]
            e2 = B2 * (cx - v2x) + C2 * (cy - v2y) + bias[2]
            covered = e0 >= 0 and e1 >= 0 and e2 >= 0
            assert covered == (winner[py, px] != 0xFFFFFFFF), (px, py)
            if not covered:
                continue
            l0 = f32(f32(B1 * (cx - v2x) + C1 * (cy - v2y)) * inv_det)
            l1 = f32(f32(B2 * (cx - v2x) + C2 * (cy - v2y)) * inv_det)
            l2 = f32(f32(f32(1.0) - l0) - l1)
            d = f32(f32(f32(l0 * z[0]) + f32(l1 * z[1])) + f32(l2 * z[2]))
            assert d.view(np.uint32) == depth[py, px].view(np.uint32), (px, py)
            b0, b1, b2 = f32(l0 * invw[0]), f32(l1 * invw[1]), f32(l2 * invw[2])
            invb = f32(f32(1.0) / f32(f32(b0 + b1) + b2))
            b0, b1 = f32(b0 * invb), f32(b1 * invb)
            b2 = f32(f32(f32(1.0) - b0) - b1)
            mix = lambda a: f32(f32(f32(b0 * a[0]) + f32(b1 * a[1])) + f32(b2 * a[2]))
            pos = [mix(P[:, k]) for k in range(3)]
            n = [mix(N[:, k]) for k in range(3)]
            dot = lambda a, b: f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))
            rs = lambda x: f32(f32(1.0) / np.sqrt(x, dtype=f32))
            w_ = rs(dot(n, n))
            n = [f32(k * w_) for k in n]
            da = dot(light, n)
            da = f32(0.0) if da < 0 else da
            diffuse = f32(f32(f32(da + f32(0.2)) * f32(3.0)) * f32(0.31830988618))
            e = [f32(eye[k] - pos[k]) for k in range(3)]
            w_ = rs(dot(e, e))
            e = [f32(k * w_) for k in e]
            hv = [f32(light[k] + e[k]) for k in range(3)]
            w_ = rs(dot(hv, hv))
            hv = [f32(k * w_) for k in hv]
            spec = f32(np.power(dot(n, hv), f32(200.0), dtype=f32) * f32(3.0))
            val = f32(diffuse + spec)
            t = f32(0.0) if val < 0 else (f32(1.0) if val > 1 else val)
            byte = int(f32(f32(t * f32(255.0)) + f32(0.5)))
            got = color[py, px]
            assert abs(int(got[0]) - byte) <= 1 and got[0] == got[1] == got[2] and got[3] == 255, (px, py, got, byte)
            checked += 1
    assert checked > 300
