"""Known-answer tests of the oracle's texture path (DESIGN.md shims 18-24; Core/Shader.h:209-244) against scalar
numpy float32 restatements written independently of it, and of the plumbing around the sampler: quad lanes, the two
differentials, per-triangle slots."""
import numpy as np
import pytest

from edxraster_b200 import scenes
from oracle import orc
from test_oracle_kats import raster_scene

f32 = np.float32
K = f32(1.0) / f32(255.0)


def to_u8(c):
    t = f32(min(max(c, f32(0.0)), f32(1.0)))
    return int(f32(f32(t * f32(255.0)) + f32(0.5)))


def np_mips(img):
    out = [img]
    while out[-1].shape[0] > 1 or out[-1].shape[1] > 1:
        s = out[-1]
        h, w = max(1, s.shape[0] >> 1), max(1, s.shape[1] >> 1)
        d = np.zeros((h, w, 4), np.uint8)
        for y in range(h):
            for x in range(w):
                x0, x1 = min(2 * x, s.shape[1] - 1), min(2 * x + 1, s.shape[1] - 1)
                y0, y1 = min(2 * y, s.shape[0] - 1), min(2 * y + 1, s.shape[0] - 1)
                for c in range(4):
                    v = f32(f32(f32(f32(s[y0, x0, c] * K) + f32(s[y0, x1, c] * K)) + f32(s[y1, x0, c] * K)) + f32(s[y1, x1, c] * K))
                    d[y, x, c] = to_u8(f32(v * f32(0.25)))
        out.append(d)
    return out


def np_texel(lv, x, y):
    h, w = lv.shape[:2]
    return (lv[y % h, x % w, :3].astype(f32) * K).astype(f32)


def np_bilinear(lv, u, v):
    h, w = lv.shape[:2]
    x, y = f32(f32(f32(u) * f32(w)) - f32(0.5)), f32(f32(f32(v) * f32(h)) - f32(0.5))
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = f32(x - x0), f32(y - y0)
    ix, iy = int(x0), int(y0)
    gx, gy = f32(f32(1) - fx), f32(f32(1) - fy)
    c00, c10, c01, c11 = np_texel(lv, ix, iy), np_texel(lv, ix + 1, iy), np_texel(lv, ix, iy + 1), np_texel(lv, ix + 1, iy + 1)
    return (((f32(gx * gy) * c00 + f32(fx * gy) * c10).astype(f32) + f32(gx * fy) * c01).astype(f32) + f32(fx * fy) * c11).astype(f32)


def np_trilinear(mips, u, v, width):
    L = len(mips)
    level = f32(f32(L - 1) + np.log2(max(f32(width), f32(1e-8)), dtype=f32))
    if not level >= 0:
        return np_bilinear(mips[0], u, v)
    if level >= L - 1:
        return np_texel(mips[-1], 0, 0)
    i = int(np.floor(level))
    d = f32(level - f32(i))
    return (f32(f32(1) - d) * np_bilinear(mips[i], u, v) + d * np_bilinear(mips[i + 1], u, v)).astype(f32)


@pytest.fixture(scope="module")
def tex():
    o = orc.Oracle(32, 32, 1)
    imgs = [scenes.noise_texture(16, 8, 1), scenes.noise_texture(37, 21, 2, cell=3), scenes.noise_texture(1, 5, 3, cell=1)]
    o.set_textures([("image", i) for i in imgs] + [("constant", (0.25, 0.5, 0.75))])
    return o, imgs


def test_mip_chain_is_the_2x2_box_filter_down_to_one_texel(tex):
    o, imgs = tex
    for slot, img in enumerate(imgs):
        got, want = o.tex_mips(slot), np_mips(img)
        assert [m.shape for m in got] == [m.shape for m in want]
        assert got[-1].shape[:2] == (1, 1)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)


def test_constant_texture_ignores_everything(tex):
    o, _ = tex
    for f in range(6):
        np.testing.assert_array_equal(o.tex_sample(3, f, 12.3, -4.0, (1, 2), (3, 4)), np.array([0.25, 0.5, 0.75], f32))


def test_nearest_and_bilinear_with_repeat_addressing(tex):
    o, imgs = tex
    rng = np.random.default_rng(0)
    for slot, img in enumerate(imgs):
        for _ in range(200):
            u, v = f32(rng.uniform(-3, 3)), f32(rng.uniform(-3, 3))
            h, w = img.shape[:2]
            want = np_texel(img, int(np.floor(f32(u * f32(w)))), int(np.floor(f32(v * f32(h)))))
            np.testing.assert_array_equal(o.tex_sample(slot, 0, u, v), want)
            np.testing.assert_array_equal(o.tex_sample(slot, 1, u, v), np_bilinear(img, u, v))
    # texel centres reproduce the texel, the corner between four texels their mean
    img = imgs[0]
    np.testing.assert_array_equal(o.tex_sample(0, 1, 2.5 / 16, 3.5 / 8), np_texel(img, 2, 3))
    np.testing.assert_allclose(o.tex_sample(0, 1, 3.0 / 16, 4.0 / 8), img[3:5, 2:4, :3].astype(np.float64).mean(axis=(0, 1)) / 255, atol=1e-6)


def test_trilinear_level_selection_and_blend(tex):
    o, imgs = tex
    mips = np_mips(imgs[0])                  # 16x8: 5 levels
    rng = np.random.default_rng(1)
    for _ in range(300):
        u, v = f32(rng.uniform(-1, 2)), f32(rng.uniform(-1, 2))
        d0 = (f32(rng.normal() * 0.1), f32(rng.normal() * 0.1))
        d1 = (f32(rng.normal() * 0.1), f32(rng.normal() * 0.1))
        width = f32(f32(2) * max(abs(d0[0]), abs(d0[1]), abs(d1[0]), abs(d1[1])))
        np.testing.assert_allclose(o.tex_sample(0, 2, u, v, d0, d1), np_trilinear(mips, u, v, width), atol=2e-6)
    # footprint of exactly one level-1 texel: width 2/16 -> level 4 + log2(1/8) = 1
    np.testing.assert_allclose(o.tex_sample(0, 2, 0.3, 0.6, (1 / 16, 0), (0, 0)), np_bilinear(mips[1], f32(0.3), f32(0.6)), atol=1e-6)
    # magnification -> level 0 bilinear; a huge footprint -> the 1x1 level
    np.testing.assert_array_equal(o.tex_sample(0, 2, 0.3, 0.6, (1e-4, 0), (0, 1e-4)), np_bilinear(mips[0], f32(0.3), f32(0.6)))
    np.testing.assert_array_equal(o.tex_sample(0, 2, 0.3, 0.6, (5, 0), (0, 5)), np_texel(mips[-1], 0, 0))


@pytest.mark.parametrize("filt,N", [(3, 4), (4, 8), (5, 16)])
def test_anisotropic_taps_along_the_major_axis(tex, filt, N):
    o, imgs = tex
    mips = np_mips(imgs[1])
    rng = np.random.default_rng(filt)
    for _ in range(150):
        u, v = f32(rng.uniform(0, 1)), f32(rng.uniform(0, 1))
        d0 = (f32(rng.normal() * 0.05), f32(rng.normal() * 0.05))
        d1 = (f32(rng.normal() * 0.01), f32(rng.normal() * 0.01))
        l0 = np.sqrt(f32(f32(d0[0] * d0[0]) + f32(d0[1] * d0[1])), dtype=f32)
        l1 = np.sqrt(f32(f32(d1[0] * d1[0]) + f32(d1[1] * d1[1])), dtype=f32)
        (lmaj, lmin, m) = (l0, l1, d0) if l0 >= l1 else (l1, l0, d1)
        n = N if f32(lmin * f32(N)) <= lmaj else min(N, max(1, int(np.ceil(f32(lmaj / lmin)))))
        width = f32(f32(2) * f32(lmaj / f32(n)))
        acc = np.zeros(3, f32)
        for i in range(n):
            s = f32(f32(f32(f32(i) + f32(0.5)) / f32(n)) - f32(0.5))
            acc = (acc + np_trilinear(mips, f32(u + f32(m[0] * s)), f32(v + f32(m[1] * s)), width)).astype(f32)
        np.testing.assert_allclose(o.tex_sample(1, filt, u, v, d0, d1), (acc * f32(f32(1) / f32(n))).astype(f32), atol=3e-6)
    # an isotropic footprint takes one tap: the trilinear sample of width 2 * length
    iso = o.tex_sample(1, filt, 0.4, 0.7, (0.03, 0), (0, 0.03))
    np.testing.assert_array_equal(iso, o.tex_sample(1, 2, 0.4, 0.7, (0.03, 0), (0, 0)))


@pytest.mark.parametrize("filt", [1, 2, 5])
def test_quad_lanes_differentials_and_slots_reach_the_sampler(filt):
    # Two screen-space triangles with texcoords affine in the pixel position: u = a x + b y + c, v = d x + e y + f.
    # Lane 1 of a quad is x + 1, lane 2 is y + 1 (Rasterizer.h:23), so every pixel must be sampled with the
    # differentials (a, d) and (b, e) at its own centre, from the slot of its own triangle (Shader.h:228-241).
    W, H = 48, 40
    tris = [[[2, 3], [44, 5], [6, 37]], [[44, 5], [45, 38], [6, 37]]]
    sc = raster_scene(tris, [0.5, 0.5], W, H)
    a, b, c, d, e, f = 0.031, 0.007, 0.11, -0.004, 0.052, 0.3
    px = (sc.vertices[:, 0].astype(np.float64) + 1.0) * W / 2.0
    py = (1.0 - sc.vertices[:, 1].astype(np.float64)) * H / 2.0
    v = sc.vertices.copy()
    v[:, 6], v[:, 7] = a * px + b * py + c, d * px + e * py + f
    imgs = [scenes.noise_texture(32, 16, 4), scenes.noise_texture(8, 8, 5, cell=2)]
    o = orc.Oracle(W, H, 1)
    o.set_transform(sc.mv, sc.proj, sc.raster)
    o.set_shader(scenes.SHADER_LAMBERT_ALBEDO)
    o.set_textures([("image", imgs[0]), ("image", imgs[1])], np.array([1, 0], np.uint32))
    o.set_texture_filter(filt)
    o.render(v, sc.indices)
    color, winner = o.color()[::-1], o.winner()[::-1]
    # Shader.h:256-264 with the normal (0, 0, -1) of raster_scene and the light (1, 1, -1) / sqrt(3)
    diffuse = f32(f32(f32(f32(1.0) / np.sqrt(f32(3.0))) + f32(0.2)) * f32(3.0)) * f32(0.31830988618)
    checked = 0
    for y in range(H):
        for x in range(W):
            if winner[y, x] == 0xFFFFFFFF:
                continue
            slot = [1, 0][winner[y, x] >> 3]
            uu, vv = a * (x + 0.5) + b * (y + 0.5) + c, d * (x + 0.5) + e * (y + 0.5) + f
            alb = o.tex_sample(slot, filt, uu, vv, (a, d), (b, e))
            want = [to_u8(f32(diffuse * f32(ch))) for ch in alb]
            got = color[y, x]
            assert all(abs(int(got[k]) - want[k]) <= 1 for k in range(3)) and got[3] == 255, (x, y, got, want)
            checked += 1
    assert checked > 1200
    assert len({tuple(px) for px in color.reshape(-1, 4)}) > 300        # it really is textured
