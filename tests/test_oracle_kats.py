"""Known-answer tests that pin the oracle (CPU, no GPU).

The reference ships no tests or golden vectors, so these expectations are derived by
hand from its formulas; each test cites the lines it follows (relative to /root/reference/EDXRaster/).
"""
import numpy as np
import pytest

from edxraster_b200 import camera as cam
from edxraster_b200 import scenes
from oracle import orc

I4 = np.eye(4, dtype=np.float32)


def raster_scene(tris_px, z, width=16, height=16):
    """Triangles given directly in raster pixels (y down). MV = P = identity, w = 1."""
    tris_px = np.asarray(tris_px, np.float64).reshape(-1, 3, 2)
    z = np.broadcast_to(np.asarray(z, np.float64).reshape(-1, 1) if np.ndim(z) == 1 else np.asarray(z, np.float64), tris_px.shape[:2])
    pos = np.empty(tris_px.shape[:2] + (3,))
    pos[..., 0] = 2.0 * tris_px[..., 0] / width - 1.0
    pos[..., 1] = 1.0 - 2.0 * tris_px[..., 1] / height
    pos[..., 2] = z
    n = pos.shape[0] * 3
    v = np.zeros((n, 8), np.float32)
    v[:, 0:3] = pos.reshape(-1, 3)
    v[:, 5] = -1.0
    return scenes.Scene(name="kat", width=width, height=height, vertices=v,
                        indices=np.arange(n, dtype=np.uint32).reshape(-1, 3), mv=I4, proj=I4,
                        raster=cam.raster_matrix(width, height), shader=0)


def run(sc, **kw):
    o = orc.Oracle(sc.width, sc.height, kw.get("threads", 1))
    o.set_transform(sc.mv, sc.proj, sc.raster)
    o.set_shader(sc.shader)
    o.set_hierarchical(kw.get("hierarchical", True))
    o.render(sc.vertices, sc.indices)
    return o


def covered(o):
    """covered pixel mask in raster orientation (row 0 = top)"""
    return (o.winner() != 0xFFFFFFFF)[::-1]


def test_snap_truncates_toward_zero():
    # RasterTriangle.h:35-40: int = float * 16.0 truncates (C++ conversion)
    assert orc.snap(1.0) == 16
    assert orc.snap(0.99) == 15
    assert orc.snap(1.0 / 16.0) == 1
    assert orc.snap(-0.03) == 0            # -0.48 -> 0, not -1
    assert orc.snap(-1.99 / 16.0) == -1
    assert orc.snap(100.53125) == 1608     # exact .5 sub-pixel stays
    assert orc.snap(float("nan")) == -2**31
    assert orc.snap(1e30) == -2**31


def test_clip_codes():
    # Clipper.h:48-68: LEFT 1, RIGHT 2, BOTTOM 4, TOP 8, NEAR 16, FAR 32; D3D volume 0 <= z <= w
    assert orc.clip_code(0, 0, 0.5, 1) == 0
    assert orc.clip_code(-1, 1, 0, 1) == 0               # on the planes = inside
    assert orc.clip_code(-1.5, 0, 0.5, 1) == 1
    assert orc.clip_code(1.5, 0, 0.5, 1) == 2
    assert orc.clip_code(0, -1.5, 0.5, 1) == 4
    assert orc.clip_code(0, 1.5, 0.5, 1) == 8
    assert orc.clip_code(0, 0, -0.1, 1) == 16
    assert orc.clip_code(0, 0, 1.1, 1) == 32
    assert orc.clip_code(2, 2, -1, 1) == 2 | 8 | 16
    assert orc.clip_code(0, 0, 0, -1) == (1 | 2 | 4 | 8 | 32)   # w < 0: outside both planes of each pair


def test_clip_near_plane_single_vertex_behind():
    # Clipper.h:273-278: t = z0 / (z0 - z1), new z snapped to exactly 0; order: kept vertex, cut, cut
    tri = [[0, 0, 0.5, 1], [1, 0, -0.5, 1], [0, 1, 0.5, 1]]
    pos, wt = orc.clip_triangle(tri)
    assert pos.shape[0] == 4
    assert (pos[:, 2] >= 0).all()
    assert np.count_nonzero(pos[:, 2] == 0.0) == 2
    # polygon order produced by Clipper.h:196-229 starting at edge (v0,v1): cut(v0,v1), cut(v1,v2), v2, v0
    np.testing.assert_allclose(pos[0], [0.5, 0, 0, 1])
    np.testing.assert_allclose(pos[1], [0.5, 0.5, 0, 1])
    np.testing.assert_array_equal(wt[2], [0, 0, 1])
    np.testing.assert_array_equal(wt[3], [1, 0, 0])
    np.testing.assert_allclose(wt[0], [0.5, 0.5, 0])


@pytest.mark.parametrize("axis,sign,bit", [(0, -1, "x=-w"), (0, 1, "x=w"), (1, -1, "y=-w"), (1, 1, "y=w")])
def test_clip_side_planes_snap_exactly(axis, sign, bit):
    # Clipper.h:238-264: the clipped coordinate is overwritten with +-w, not left to rounding
    tri = np.array([[0, 0, 0.5, 1.0], [0.3, 0.2, 0.5, 1.3], [0.1, 0.4, 0.5, 0.9]], np.float32)
    tri[1, axis] = sign * 2.7
    pos, wt = orc.clip_triangle(tri)
    assert pos.shape[0] == 4
    new = [p for p, w in zip(pos, wt) if not (w == 1.0).any()]
    assert len(new) == 2
    for p in new:
        assert p[axis] == sign * p[3]


def test_clip_rejects_when_all_outside_one_plane_and_drops_w_le_zero():
    assert orc.clip_triangle([[2, 0, 0.5, 1], [3, 0, 0.5, 1], [2, 1, 0.5, 1]])[0].shape[0] == 0     # Clipper.h:109
    assert orc.clip_triangle([[0, 0, 0.5, 1], [0.1, 0, 0.5, 1], [0, 0.1, 0.5, 1]])[0] is None       # not clipped at all
    # a vertex with w <= 0 that survives the enabled planes empties the polygon (Clipper.h:280-287)
    pos, _ = orc.clip_triangle([[0, 0, 0.0, 0.0], [0.5, 0, 0.5, 1], [0, 2.0, 0.5, 1]])
    assert pos.shape[0] == 0


def test_fan_order_and_prim_ids():
    # Clipper.h:156-170: fan (0, k-1, k); our prim id = triangle * 8 + (k - 2)
    # first triangle has one vertex left of the screen -> clipped to a quad -> 2 fan triangles
    # (front-facing: raster cross(v0 - v2, v1 - v2) > 0)
    sc = raster_scene([[[-8, 8], [6, 2], [6, 14]], [[2, 2], [14, 2], [2, 14]]], [0.5, 0.5])
    o = run(sc)
    ints, _ = o.raster_tris()
    assert ints[:, 0].tolist() == [0, 1, 8]
    # fan triangles share polygon vertex 0 (Clipper.h:158) and every clipped x is exactly 0
    assert (ints[0, 1:3] == ints[1, 1:3]).all()
    assert ints[:2, 1::2].min() == 0


def test_fill_rule_shared_edge_covers_every_centre_once():
    # Two triangles sharing the diagonal of a square whose corners are pixel CENTRES.
    # Top-left rule (RasterTriangle.h:296-299 as a -1 bias): each centre on the shared edge belongs to
    # exactly one triangle, and of the square's border only the top and left edges are drawn.
    c = lambda i: i + 0.5
    a, b = c(2), c(10)
    quad = [[[a, a], [b, a], [a, b]], [[b, a], [b, b], [a, b]]]
    sc = raster_scene(quad, [0.5, 0.5])
    o = run(sc)
    st = o.stats()
    cov = covered(o)
    # pixels x,y in [2, 9] are covered (top/left inclusive, bottom/right exclusive) -> 8 x 8
    expect = np.zeros((16, 16), bool)
    expect[2:10, 2:10] = True
    np.testing.assert_array_equal(cov, expect)
    assert st["covered_samples"] == 64          # no centre was covered twice


def test_fill_rule_vertex_on_centre_and_horizontal_edges():
    c = lambda i: i + 0.5
    # top edge horizontal through centres of row 3 -> drawn; bottom vertex on centre (6,9) -> not drawn
    sc = raster_scene([[[c(3), c(3)], [c(9), c(3)], [c(6), c(9)]]], [0.5])
    cov = covered(run(sc))
    assert cov[3, 3] and cov[3, 8] and not cov[3, 9]        # top edge: left end in, right end out
    assert not cov[9, 6]                                     # bottom apex lies on a right/bottom edge
    # flipped: bottom edge horizontal through row 9 -> not drawn; top apex on centre (6,3) is drawn
    sc = raster_scene([[[c(6), c(3)], [c(9), c(9)], [c(3), c(9)]]], [0.5])
    cov = covered(run(sc))
    assert not cov[9, 3:10].any()
    assert cov[8, 6]


def test_backfacing_and_degenerate_are_culled():
    # RasterTriangle.h:49-51: det <= 0 -> culled
    front = [[2.5, 2.5], [9.5, 2.5], [2.5, 9.5]]
    back = [front[1], front[0], front[2]]
    degenerate = [[2.5, 2.5], [5.5, 5.5], [8.5, 8.5]]
    o = run(raster_scene([front], [0.5]))
    assert o.stats()["raster_tris"] == 1
    assert run(raster_scene([back], [0.5])).stats()["raster_tris"] == 0
    assert run(raster_scene([degenerate], [0.5])).stats()["raster_tris"] == 0


def test_depth_ties_later_wins_and_clear_value_rules():
    # FrameBuffer.cpp:64-65: LESS_EQUAL with immediate write; cleared to 1.0 (:103)
    t = [[2.5, 2.5], [12.5, 2.5], [2.5, 12.5]]
    o = run(raster_scene([t, t, t], [0.5, 0.25, 0.25]))
    w = o.winner()[::-1]
    assert w[4, 4] == 16                      # equal depth: the later triangle (id 2 -> prim 16) owns the pixel
    assert o.depth()[::-1][4, 4] == np.float32(0.25)
    o = run(raster_scene([t, t], [0.25, 0.5]))
    assert o.winner()[::-1][4, 4] == 0        # farther later triangle fails
    # z == 1.0 exactly passes against the clear value; z slightly above w is clipped away (FAR plane)
    o = run(raster_scene([t], [1.0]))
    assert o.winner()[::-1][4, 4] == 0 and o.depth()[::-1][4, 4] == np.float32(1.0)
    o = run(raster_scene([t], [1.5]))
    assert (o.winner() == 0xFFFFFFFF).all()


def test_frame_buffer_is_bottom_up_rgba():
    # FrameBuffer.cpp:41 + Main.cpp:75 (glDrawPixels): row 0 is the bottom scanline
    sc = raster_scene([[[1.5, 1.5], [6.5, 1.5], [1.5, 6.5]]], [0.5])
    sc["shader"] = 1
    o = run(sc)
    col = o.color()
    assert col.shape == (16, 16, 4)
    assert col[15 - 2, 2, 3] == 255 and col[2, 2, 3] == 0       # drawn near the top-left -> high row index
    assert (col[col[..., 3] == 0] == 0).all()                  # cleared pixels are all-zero (FrameBuffer.cpp:91-95)


def test_matrix_shims():
    rng = np.random.default_rng(1)
    a = (rng.random((4, 4)) + np.eye(4) * 2).astype(np.float32)
    b = rng.random((4, 4)).astype(np.float32)
    np.testing.assert_allclose(orc.mat_mul(a, b), a @ b, rtol=1e-6)
    np.testing.assert_allclose(orc.mat_inverse(a) @ a, np.eye(4), atol=1e-5)


def test_msaa_sample_positions_and_box_resolve():
    # FrameBuffer.cpp:107-191 sample tables (1/16 px around the centre), Rasterizer.h:245-270 per-sample
    # coverage, FrameBuffer.cpp:70-87 box resolve. A vertical edge through x = 5.5 (the centre column of
    # pixel 5): samples with offset.x < 0 are left of it. 4x table: (-2,-6) (6,-2) (-6,2) (2,6) -> samples 0, 2
    # are covered by the left half-plane; resolved alpha = 2/4 * 255 -> 128 (FromFloats rounding).
    sc = raster_scene([[[1.5, 1.5], [5.5, 1.5], [5.5, 12.5]], [[1.5, 1.5], [5.5, 12.5], [1.5, 12.5]]], [0.5, 0.5])
    sc["shader"] = 2
    o = orc.Oracle(sc.width, sc.height, 1)
    o.set_msaa(2)
    o.set_transform(sc.mv, sc.proj, sc.raster)
    o.set_shader(2)
    o.render(sc.vertices, sc.indices)
    assert o.samples == 4
    cov = [(o.winner(k) != 0xFFFFFFFF)[::-1] for k in range(4)]
    assert cov[0][6, 5] and cov[2][6, 5] and not cov[1][6, 5] and not cov[3][6, 5]
    assert all(c[6, 3] for c in cov)                      # interior pixel: every sample
    assert not any(c[6, 6] for c in cov)                  # pixel right of the edge: none
    col = o.color()[::-1]
    assert col[6, 3, 3] == 255 and col[6, 5, 3] == 128 and col[6, 6, 3] == 0
    # colour is shaded once per fragment at the pixel centre, so the half-covered pixel is half the interior colour
    assert abs(int(col[6, 5, 0]) - int(round(col[6, 3, 0] / 2))) <= 1
