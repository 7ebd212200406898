"""The oracle against THE REFERENCE ITSELF (CPU, no GPU).

oracle/_ref/libref.so is the reference's own Core/*.cpp, Core/*.h and Utils/* compiled unmodified by g++ against the
EDXUtil stand-in in oracle/_ref_shim (recipe: oracle/Makefile). These tests render the same inputs through it and
through the restatement (oracle/edx_oracle.cpp) and require identical bits: clip-space vertices, snapped set-up
records, per-pixel (and per-sample) depth and owner, and the RGBA8 back buffer byte for byte (GetBackBuffer,
Renderer.cpp:360). That is what lets the `-m gpu` suite, which compares the CUDA path with the restatement at sizes
the reference build cannot reach in seconds, speak for the reference.

/root/reference exists only in the build container; on a box without the prebuilt library the module is skipped.

Frame sizes: the reference writes out of bounds when a partial tile is narrower / shorter than 16 px or a dimension is
odd (Rasterizer.h:92-95, FrameBuffer.cpp:41), so every size here has W, H even and
W mod 32, H mod 32 in {0, 16..31}.
"""
import copy

import numpy as np
import pytest

from edxraster_b200 import camera as cam
from edxraster_b200 import scenes
import parity
from test_oracle_kats import raster_scene

pytestmark = pytest.mark.skipif(not parity.reference_available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")


def assert_same(sc, **kw):
    ref = parity.render_reference(sc, threads=kw.get("threads", 0), shader=kw.get("shader"), hierarchical=kw.get("hierarchical", True), msaa=kw.get("msaa", 0))
    got = parity.render_oracle(sc, threads=kw.get("oracle_threads", 0), shader=kw.get("shader"), hierarchical=kw.get("hierarchical", True), msaa=kw.get("msaa", 0))
    rep = parity.compare_reference(ref, got)
    assert parity.is_reference_parity(rep), rep
    return ref, got, rep


CONFIGS = {
    "C1": lambda: scenes.config1(width=640, height=368, slices=64, stacks=64),
    "C2": lambda: scenes.config2(width=640, height=368, num_tris=40000),
    "C3": lambda: scenes.config3(width=640, height=368, num_tris=120),
    "C4": lambda: scenes.config4(width=640, height=368, quads_x=240, quads_z=192),
    "C4_yaw": lambda: scenes.config4(width=496, height=272, quads_x=160, quads_z=128, yaw=2.1),
    "C5_view": lambda: _c5_view(),
}


def _c5_view():
    base = scenes.by_name("C4", 0.01)
    sc = copy.copy(base)
    sc.mv, sc.proj, sc.raster = scenes.config5_views(base, 7)[3]
    return sc


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_configs_identical_to_the_reference(name):
    sc = CONFIGS[name]()
    ref, got, rep = assert_same(sc)
    assert rep["tri_count"][0] > 0
    if sc.shader in (1, 3):
        assert rep["color_diff_pixels"] == 0 and (ref["color"][..., 3] == 255).sum() > 100


def test_full_size_c1_identical_to_the_reference():
    # BASELINE.json configs[0], the reference's own CPU-runnable case, at its full size (1280x720, 20 000 triangles)
    assert_same(scenes.config1())


@pytest.mark.parametrize("shader", [1, 3])
def test_both_reference_shaders(shader):
    _, _, rep = assert_same(scenes.config1(width=320, height=208, slices=40, stacks=40), shader=shader)
    assert rep["color_diff_pixels"] == 0


@pytest.mark.parametrize("threads", [1, 2, 5, 12])
def test_reference_is_invariant_under_its_core_count(threads):
    # Tile::triangleRefs[12] (Tile.h:34) caps the reference at 12 cores; results do not depend on the count
    sc = scenes.config4(width=320, height=208, quads_x=120, quads_z=96)
    ref, _, _ = assert_same(sc, threads=threads, oracle_threads=3)
    assert ref["threads"] == threads


def test_hierarchical_off_is_the_same_image():
    sc = scenes.config3(width=320, height=208, num_tris=40)
    a, _, _ = assert_same(sc, hierarchical=False)
    b, _, _ = assert_same(sc, hierarchical=True)
    np.testing.assert_array_equal(a["color"], b["color"])
    np.testing.assert_array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))


# ---- the hand-derived known answers of tests/test_oracle_kats.py hold for the reference's own code ----

def _cov(ref):
    return (ref["winner_ord"] != 0xFFFFFFFF)[::-1]


def test_fill_rule_of_the_reference_is_top_left():
    # RasterTriangle.h:296-299 compiled as written (BoolSSE -> IntSSE is a bit-cast: DESIGN.md shim 5)
    c = lambda i: i + 0.5
    a, b = c(2), c(10)
    sc = raster_scene([[[a, a], [b, a], [a, b]], [[b, a], [b, b], [a, b]]], [0.5, 0.5], 32, 32)
    ref, _, _ = assert_same(sc)
    expect = np.zeros((32, 32), bool)
    expect[2:10, 2:10] = True
    np.testing.assert_array_equal(_cov(ref), expect)
    sc = raster_scene([[[c(3), c(3)], [c(9), c(3)], [c(6), c(9)]]], [0.5], 32, 32)
    cov = _cov(assert_same(sc)[0])
    assert cov[3, 3] and cov[3, 8] and not cov[3, 9] and not cov[9, 6]


def test_depth_ties_and_clear_value_in_the_reference():
    t = [[2.5, 2.5], [12.5, 2.5], [2.5, 12.5]]
    ref, _, _ = assert_same(raster_scene([t, t, t], [0.5, 0.25, 0.25], 32, 32))
    assert ref["winner_ord"][::-1][4, 4] == 2 and ref["depth"][::-1][4, 4] == np.float32(0.25)
    ref, _, _ = assert_same(raster_scene([t], [1.0], 32, 32))
    assert ref["depth"][::-1][4, 4] == np.float32(1.0) and ref["winner_ord"][::-1][4, 4] == 0
    ref, _, _ = assert_same(raster_scene([t], [1.5], 32, 32))
    assert (ref["winner_ord"] == 0xFFFFFFFF).all()


def test_fan_order_culls_and_bottom_up_buffer():
    sc = raster_scene([[[-8, 8], [6, 2], [6, 14]], [[2, 2], [14, 2], [2, 14]], [[9.5, 2.5], [2.5, 2.5], [2.5, 9.5]],
                       [[2.5, 2.5], [5.5, 5.5], [8.5, 8.5]]], [0.5, 0.5, 0.5, 0.5], 32, 32)
    sc["shader"] = 1
    ref, got, rep = assert_same(sc)
    assert rep["tri_count"] == (3, 3)                        # quad fan (2) + 1; back-facing and degenerate culled
    ints, _ = ref["tris"]
    assert (ints[0, 0:2] == ints[1, 0:2]).all() and ints[:2, 0::2].min() == 0
    col = ref["color"]
    assert col[31 - 4, 4, 3] == 255 and col[4, 20, 3] == 0   # row 0 is the bottom scanline (FrameBuffer.cpp:41)


def test_every_clip_plane_and_the_w_le_zero_drop():
    rng = np.random.default_rng(3)
    n = 4000
    sc = scenes.config3(width=320, height=208, num_tris=n)
    v = sc.vertices.copy()
    v[:, 0:3] = (rng.random((n * 3, 3)) - 0.5) * np.array([8.0, 8.0, 6.0])
    sc["vertices"] = v
    _, _, rep = assert_same(sc)
    assert rep["tri_count"][0] > 500


# ---- MSAA (Rasterizer.h:202-300,355-415; FrameBuffer.cpp:70-87,107-191) ----

MSAA_SCENES = {
    "C1": lambda: scenes.config1(width=320, height=208, slices=40, stacks=40),
    "C3": lambda: scenes.config3(width=320, height=208, num_tris=30),
    "C4": lambda: scenes.config4(width=320, height=208, quads_x=120, quads_z=96),
}


@pytest.mark.parametrize("name", sorted(MSAA_SCENES))
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_msaa_every_sample_identical_to_the_reference(name, level):
    ref, _, rep = assert_same(MSAA_SCENES[name](), msaa=level)
    assert rep["color_diff_pixels"] == 0
    assert len(ref["samples"]) == 1 << level


def test_msaa_sample_positions_of_the_reference_lie_on_the_diagonal():
    # Rasterizer.h:247 binds `const Vector2i&` to ONE int of the offset table: sample s sits at (t[2s], t[2s])
    # (DESIGN.md shim 17). 4x: (-2,-2) (6,6) (-6,-6) (2,2). A horizontal edge through y = 6.5 with the triangle above
    # it covers exactly the samples with a negative offset (0 and 2) of the pixels in row 6; the written pairs
    # (-2,-6) (6,-2) (-6,2) (2,6) would cover samples 0 and 1 instead.
    sc = raster_scene([[[1.5, 1.5], [12.5, 1.5], [12.5, 6.5]], [[1.5, 1.5], [12.5, 6.5], [1.5, 6.5]]], [0.5, 0.5], 32, 32)
    ref = parity.render_reference(sc, msaa=2)
    cov = [(w != 0xFFFFFFFF)[::-1] for _, w in ref["samples"]]
    assert cov[0][6, 5] and cov[2][6, 5] and not cov[1][6, 5] and not cov[3][6, 5]
    assert ref["color"][::-1][6, 5, 3] == 128               # two of four samples: alpha 0.5 -> 128 (FrameBuffer.cpp:85)


# ---- textured default shader: the sampler is the stand-in's (DESIGN.md shims 19-24), everything around it the reference's ----

@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4, 5])
def test_textured_plane_every_filter(filt):
    ref, _, rep = assert_same(scenes.textured_plane(width=320, height=208, tex_filter=filt))
    assert rep["color_diff_pixels"] == 0
    assert len(np.unique(ref["color"].reshape(-1, 4), axis=0)) > 1000


@pytest.mark.parametrize("filt", [0, 2, 4])
def test_textured_sphere_slots(filt):
    assert_same(scenes.textured_sphere(width=320, height=208, tex_filter=filt))


# ---- random soups ----

def _fuzz_scene(seed):
    rng = np.random.default_rng(seed)
    w, h = int(rng.choice([160, 256, 336])), int(rng.choice([96, 192, 144]))
    n = 2500
    centre = (rng.random((n, 1, 3)) - 0.5) * np.array([6.0, 4.0, 10.0]) + np.array([0.0, 0.0, 3.0])
    size = rng.choice([0.01, 0.05, 0.3, 2.0, 12.0], size=(n, 1, 1), p=[0.3, 0.3, 0.2, 0.15, 0.05])
    pos = centre + (rng.random((n, 3, 3)) - 0.5) * size
    v = np.zeros((n * 3, 8), np.float32)
    v[:, 0:3] = pos.reshape(-1, 3)
    v[:, 3:6] = rng.normal(size=(n * 3, 3))
    v[:, 6:8] = (rng.random((n * 3, 2)) - 0.5) * float(rng.choice([1.0, 4.0, 40.0]))
    eye = (rng.random(3) - 0.5) * np.array([2.0, 2.0, 2.0]) + np.array([0.0, 0.0, -2.0])
    c = cam.Camera(eye, (0.0, 0.0, 3.0), (0.0, 1.0, 0.0), w, h, float(rng.choice([40.0, 65.0, 100.0])),
                   float(rng.choice([0.05, 0.5, 2.0])), float(rng.choice([6.0, 50.0])))
    sc = scenes.Scene(name="fuzz%d" % seed, width=w, height=h, vertices=v,
                      indices=np.arange(n * 3, dtype=np.uint32).reshape(-1, 3), mv=c.view, proj=c.proj, raster=c.raster,
                      shader=int(rng.choice([1, 3])))
    if sc.shader == 3:
        tex = []
        for k in range(int(rng.integers(1, 4))):
            if rng.random() < 0.25:
                tex.append(("constant", tuple(rng.random(3))))
            else:
                tex.append(("image", scenes.noise_texture(int(rng.choice([1, 2, 5, 16, 33, 128])), int(rng.choice([1, 3, 8, 64])), seed + k, cell=2)))
        sc["textures"], sc["tex_ids"] = tex, rng.integers(0, len(tex), n).astype(np.uint32)
        sc["tex_filter"] = int(rng.integers(0, 6))
    return sc, int(rng.choice([0, 0, 1, 2, 3, 4, 5]))


@pytest.mark.parametrize("seed", list(range(300, 312)))
def test_fuzz_random_scenes_identical_to_the_reference(seed):
    sc, msaa = _fuzz_scene(seed)
    assert_same(sc, msaa=msaa)
