"""The oracle's results must not depend on its thread count or on the hierarchical switch
(SURVEY.md §3.3 / §3.4): this is what licenses the CUDA path to re-tile and re-order the work."""
import numpy as np
import pytest

from edxraster_b200 import scenes
import parity

SCENES = {
    "C1": lambda: scenes.config1(width=640, height=360, slices=40, stacks=40),
    "C2": lambda: scenes.config2(width=640, height=360, num_tris=30000),
    "C3": lambda: scenes.config3(width=640, height=360, num_tris=50),
    "C4": lambda: scenes.config4(width=640, height=360, quads_x=200, quads_z=160),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_thread_count_and_hierarchy_do_not_change_the_frame(name):
    sc = SCENES[name]()
    base = parity.render_oracle(sc, threads=1, hierarchical=True)
    for threads, hier in ((2, True), (8, True), (1, False), (5, False)):
        other = parity.render_oracle(sc, threads=threads, hierarchical=hier)
        np.testing.assert_array_equal(base["depth"].view(np.uint32), other["depth"].view(np.uint32))
        np.testing.assert_array_equal(base["winner"], other["winner"])
        np.testing.assert_array_equal(base["color"], other["color"])
        assert base["tris"][0].tolist() == other["tris"][0].tolist()


@pytest.mark.parametrize("name", ["C1", "C3", "C4"])
def test_msaa_frames_do_not_depend_on_threads_or_hierarchy(name):
    sc = SCENES[name]()
    base = parity.render_oracle(sc, threads=1, hierarchical=True, msaa=2)
    for threads, hier in ((4, True), (3, False)):
        other = parity.render_oracle(sc, threads=threads, hierarchical=hier, msaa=2)
        np.testing.assert_array_equal(base["color"], other["color"])
        for (d0, w0), (d1, w1) in zip(base["samples"], other["samples"]):
            np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
            np.testing.assert_array_equal(w0, w1)


def test_closed_form_of_the_depth_test():
    """Per pixel: final depth = min over covering fragments, owner = LAST fragment at that depth
    (SURVEY.md §3.3). Checked by brute force against the oracle's sequential LESS_EQUAL test."""
    from test_oracle_kats import raster_scene, run
    rng = np.random.default_rng(7)
    tris, zs = [], []
    for _ in range(60):
        c = rng.random(2) * 16
        p = c + (rng.random((3, 2)) - 0.5) * 14
        a, b = p[0] - p[2], p[1] - p[2]
        if a[0] * b[1] - b[0] * a[1] < 0:
            p = p[[1, 0, 2]]
        tris.append(p)
        zs.append(np.round(rng.random() * 4) / 4 * 0.8 + 0.1)      # few distinct depths -> many ties
    sc = raster_scene(tris, zs)
    o = run(sc)
    winner, depth = o.winner()[::-1], o.depth()[::-1]
    best_d = np.full((16, 16), np.float32(1.0))
    best_i = np.full((16, 16), 0xFFFFFFFF, np.uint32)
    for i in range(len(tris)):
        single = run(raster_scene([tris[i]], [zs[i]]))
        sw = single.winner()[::-1]
        cov = sw != 0xFFFFFFFF
        d = single.depth()[::-1]
        take = cov & (d <= best_d)
        best_d[take] = d[take]
        best_i[take] = i * 8 + (sw[take] & 7)      # fan index when the triangle was clipped by the screen edge
    np.testing.assert_array_equal(winner, best_i)
    np.testing.assert_array_equal(depth, best_d)
