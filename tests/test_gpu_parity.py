"""Parity of the CUDA path against the oracle, through the C ABI (needs a B200: -m gpu).

Bar (BASELINE.json north_star): clip-space vertices, snapped setup records, per-pixel depth bits and
owning-triangle ids bit-exact; shaded colour within +-1 of 8 bits per channel.
"""
import numpy as np
import pytest

from edxraster_b200 import scenes
import parity
from test_oracle_kats import raster_scene

pytestmark = pytest.mark.gpu

REDUCED = {
    "C1": lambda: scenes.config1(),
    "C2": lambda: scenes.config2(num_tris=150000),
    "C3": lambda: scenes.config3(width=1280, height=720, num_tris=300),
    "C4": lambda: scenes.config4(quads_x=600, quads_z=480),
    "C4_yaw": lambda: scenes.config4(width=1280, height=720, quads_x=400, quads_z=320, yaw=2.1),
    "odd_bins": lambda: scenes.config1(width=1000, height=562, slices=60, stacks=60),
}


def assert_parity(sc, **kw):
    ref = parity.render_oracle(sc, shader=kw.get("shader"), msaa=kw.get("msaa", 0))
    got = parity.render_gpu(sc, **kw)
    rep = parity.compare(ref, got)
    assert parity.is_parity(rep), rep
    return ref, got, rep


@pytest.mark.parametrize("name", sorted(REDUCED))
def test_reduced_configs_bit_exact(name):
    assert_parity(REDUCED[name]())


@pytest.mark.parametrize("name", ["C1", "C3", "C4"])
@pytest.mark.parametrize("options,hier", [({"hiz": 0}, True), ({}, False), ({"small_max": 0}, True),
                                          ({"small_max": 2}, True), ({"small_max": 40}, True),
                                          ({"cluster_cull": 2, "pdl": 0}, True), ({"cluster_cull": 0, "small_max_clip": 0}, True),
                                          ({"lean_resolve": 0, "clip_carveout": 2}, True), ({"lean_resolve": 2, "clip_carveout": 1}, True),
                                          ({"graphs": 0}, True), ({"graphs": 2, "pdl": 0}, True), ({"mid_max": 0}, True), ({"mid_max": 0, "small_max": 32}, True), ({"mid_max": 16}, True),
                                          ({"bin_min": -1}, True), ({"bin_min": -1, "mid_max": 0, "small_max": 2}, True), ({"bin_min": -1, "hiz": 0, "mid_max": 0}, False),
                                          ({"mid_max": 512, "small_max": 0, "small_max_clip": 0}, True), ({"mid_max": 200, "small_max": 3}, False)])
def test_tuning_knobs_never_change_the_image(name, options, hier):
    # routing (direct vs tile path), hierarchical Z and the 8x8 block tests are pure optimisations
    assert_parity(REDUCED[name](), options=options, hierarchical=hier, stages=False)


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C4_yaw", "odd_bins"])
@pytest.mark.parametrize("front_end,cull", [(1, 2), (2, 2), (2, 0), (1, 0)])
def test_list_front_end_is_bit_identical(name, front_end, cull):
    # front_end 1: device-side cluster cull into a work list + persistent geometry kernel; 2: plus the per-vertex
    # stage (Renderer::VertexProcessing, Renderer.cpp:120-127) whose records the geometry kernel and the resolve gather.
    # Large meshes select them automatically; forced here on the reduced configs, with and without cluster culling.
    assert_parity(REDUCED[name](), options={"front_end": front_end, "cluster_cull": cull})


@pytest.mark.parametrize("front_end", [1, 2])
def test_list_front_end_with_msaa_textures_and_frames_in_sequence(front_end):
    from edxraster_b200 import renderer as R
    r = R.Renderer(0)
    opts = {"front_end": front_end, "cluster_cull": 2}
    sc = scenes.config4(width=640, height=360, quads_x=200, quads_z=160)
    for msaa in (0, 2, 0):
        ref = parity.render_oracle(sc, msaa=msaa)
        got = parity.render_gpu(sc, msaa=msaa, stages=False, renderer=r, options=opts)
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), (msaa, rep)
    r.close()
    assert_parity(scenes.textured_plane(tex_filter=3), options=opts)
    assert_parity(scenes.textured_sphere(tex_filter=2), options=opts, stages=False)


@pytest.mark.parametrize("shader", [0, 1, 2, 3])
def test_every_shader(shader):
    sc = scenes.config1(width=640, height=360, slices=64, stacks=64)
    assert_parity(sc, shader=shader, stages=False)


MSAA_SCENES = {
    "C1": lambda: scenes.config1(width=640, height=360, slices=64, stacks=64),
    "C2": lambda: scenes.config2(width=640, height=360, num_tris=40000),
    "C3": lambda: scenes.config3(width=640, height=360, num_tris=60),
    "C4": lambda: scenes.config4(width=640, height=360, quads_x=300, quads_z=240),
}


@pytest.mark.parametrize("name", sorted(MSAA_SCENES))
@pytest.mark.parametrize("level", [1, 2, 4])
def test_msaa_every_sample_bit_exact(name, level):
    # SetMSAAMode (Renderer.cpp:94-98): per-sample coverage, depth and owner bit-exact, resolved colour +-1
    sc = MSAA_SCENES[name]()
    shader = 1 if sc.shader == 0 and name != "C2" else sc.shader
    assert_parity(sc, msaa=level, shader=shader, stages=False)


@pytest.mark.parametrize("name,level", [("C3", 2), ("C4", 1), ("C1", 3)])
def test_msaa_with_per_bin_lists_and_without_the_mid_path(name, level):
    # the per-bin lists (stage a7) are shared by the sample CTAs of a bin; forced here on short lists, with every
    # non-small triangle on the tile path, and in a second frame of the same mesh (hints from the first frame apply)
    from edxraster_b200 import renderer as R
    sc = MSAA_SCENES[name]()
    shader = 1 if sc.shader == 0 else sc.shader
    ref = parity.render_oracle(sc, msaa=level, shader=shader)
    r = R.Renderer(0)
    for _ in range(2):
        got = parity.render_gpu(sc, msaa=level, shader=shader, stages=False, renderer=r, options={"bin_min": -1, "mid_max": 0})
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), rep
    assert got["stats"]["bin_pairs"] > 0
    r.close()


def test_msaa_32x_and_mode_switches():
    from edxraster_b200 import renderer as R
    sc = scenes.config1(width=320, height=200, slices=40, stacks=40)
    r = R.Renderer(0)
    for level in (5, 0, 3, 3, 0):
        ref = parity.render_oracle(sc, msaa=level)
        got = parity.render_gpu(sc, msaa=level, stages=False, renderer=r)
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), (level, rep)
    r.close()


def test_full_size_c1():
    assert_parity(scenes.config1())


def test_full_size_c2():
    _, got, _ = assert_parity(scenes.config2())
    assert got["stats"]["binned_tris"] == 0


def test_full_size_c3():
    _, got, _ = assert_parity(scenes.config3())
    assert got["stats"]["clipped_tris"] == 2000


def test_full_size_c4():
    sc = scenes.config4()
    ref = parity.render_oracle(sc)
    for fe in (-1, 0, 1):                                  # auto (= 2 for this mesh), per-cluster CTAs, work list only
        got = parity.render_gpu(sc, stages=False, options={"front_end": fe})
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), (fe, rep)
        assert got["stats"]["clipped_tris"] > 1000


def test_full_size_c5_views():
    # BASELINE.json configs[4]: views of the 10M-triangle mesh at full size, through a FrameRing as the farm renders them
    import copy
    from edxraster_b200 import renderer as R
    base = scenes.config4()
    views = scenes.config5_views(base, 256)
    pick = [5, 67, 130, 201, 255]
    ring = R.FrameRing(0, depth=3)
    ring.Initialize(base.width, base.height)
    ring.SetPixelShader(base.shader)
    mesh = ring.CreateMesh(base.vertices, base.indices)
    for lane in ring.lanes:
        lane.SetCaptureIds(True)
    tickets = [ring.Submit(mesh, *views[v]) for v in pick[:3]]
    got = {}
    for n, v in enumerate(pick):
        t = tickets[n]
        lane = ring.lanes[t % 3]
        got[v] = {"color": ring.GetBackBuffer(t).copy(), "depth": ring.GetDepthBuffer(t), "winner": lane.GetWinnerIds(), "derived": lane.DerivedState()}
        if n + 3 < len(pick):
            tickets.append(ring.Submit(mesh, *views[pick[n + 3]]))
    for v in pick:
        sc = copy.copy(base)
        sc.mv, sc.proj, sc.raster = views[v]
        rep = parity.compare(parity.render_oracle(sc), got[v])
        assert parity.is_parity(rep), (v, rep)
    assert len({g["depth"].tobytes() for g in got.values()}) == len(pick)
    mesh.Release()
    ring.close()


def test_edge_cases():
    from edxraster_b200 import renderer as R
    r = R.Renderer(0)
    # empty mesh: cleared frame (FrameBuffer.cpp:89-105)
    sc = raster_scene(np.zeros((0, 3, 2)), np.zeros((0,)), 64, 48)
    sc["shader"] = 1
    got = parity.render_gpu(sc, stages=False, renderer=r)
    assert (got["depth"] == 1.0).all() and (got["color"] == 0).all() and (got["winner"] == 0xFFFFFFFF).all()
    # everything off screen / behind the camera / degenerate / back-facing
    far_away = [[[200, 200], [300, 200], [200, 300]], [[-50, -50], [-10, -50], [-50, -10]]]
    sc = raster_scene(far_away, [0.5, 0.5], 64, 48)
    ref, got, _ = assert_parity(sc, renderer=r)
    assert (got["winner"] == 0xFFFFFFFF).all()
    tiny = raster_scene([[[0.2, 0.2], [1.9, 0.3], [0.4, 1.8]]], [0.3], 2, 2)
    assert_parity(tiny, renderer=r)
    # sub-pixel slivers that touch no centre, and a triangle covering exactly one centre
    sc = raster_scene([[[3.1, 3.1], [3.4, 3.1], [3.1, 3.4]], [[5.25, 5.25], [5.9, 5.25], [5.25, 5.9]]], [0.5, 0.5], 64, 48)
    ref, got, _ = assert_parity(sc, renderer=r)
    assert int((got["winner"] != 0xFFFFFFFF).sum()) == 1
    r.close()


def test_w_le_zero_and_huge_triangles_through_the_clipper():
    rng = np.random.default_rng(3)
    # random triangles in clip space with w of both signs: exercises every plane and the w <= 0 drop
    n = 4000
    sc = scenes.config3(width=320, height=200, num_tris=n)
    v = sc.vertices.copy()
    v[:, 0:3] = (rng.random((n * 3, 3)) - 0.5) * np.array([8.0, 8.0, 6.0])
    sc["vertices"] = v
    assert_parity(sc)


def test_repeated_frames_are_identical_and_state_switches_are_clean():
    from edxraster_b200 import renderer as R
    r = R.Renderer(0)
    a, b = scenes.config1(width=640, height=360), scenes.config4(width=640, height=360, quads_x=200, quads_z=160)
    r.Initialize(640, 360)
    r.SetCaptureIds(True)
    ma, mb = r.CreateMesh(a.vertices, a.indices), r.CreateMesh(b.vertices, b.indices)
    frames = []
    for sc, m, shader in ((a, ma, 1), (b, mb, 1), (a, ma, 0), (a, ma, 1), (b, mb, 1)):
        r.SetTransform(sc.mv, sc.proj, sc.raster)
        r.SetPixelShader(shader)
        r.RenderMesh(m)
        frames.append((r.GetBackBuffer().copy(), r.GetDepthBuffer(), r.GetWinnerIds()))
    # the self-cleaning key buffer leaves nothing behind: frame 3 == frame 0, frame 4 == frame 1
    for i, j in ((0, 3), (1, 4)):
        for x, y in zip(frames[i], frames[j]):
            np.testing.assert_array_equal(x, y)
    assert (frames[2][0] == 0).all()                      # depth-only frame leaves colour cleared
    np.testing.assert_array_equal(frames[2][1], frames[0][1])
    ref = parity.render_oracle(b)
    np.testing.assert_array_equal(ref["winner"], frames[4][2])
    # Resize (Renderer.cpp:64-83) rebuilds the buffers
    r.Resize(320, 200)
    c = scenes.config1(width=320, height=200)
    r.SetTransform(c.mv, c.proj, c.raster)
    r.RenderMesh(ma)
    np.testing.assert_array_equal(parity.render_oracle(c)["winner"], r.GetWinnerIds())
    r.close()


def test_queue_regrow_path():
    # > 65,536 triangles on the tile path overflow its initial queue: the frame is re-run after growing
    rng = np.random.default_rng(5)
    n = 120000
    c = rng.random((n, 1, 2)) * np.array([1280, 720])
    p = c + (rng.random((n, 3, 2)) - 0.5) * 90
    a, b = p[:, 0] - p[:, 2], p[:, 1] - p[:, 2]
    flip = (a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]) < 0
    p[flip, 0], p[flip, 1] = p[flip, 1].copy(), p[flip, 0].copy()
    sc = raster_scene(p, 0.1 + 0.8 * rng.random((n, 3)), 1280, 720)
    _, got, _ = assert_parity(sc, stages=False, options={"mid_max": 0, "small_max": 32})
    assert got["stats"]["binned_tris"] > 83000 and got["stats"]["regrow_count"] >= 1
    # the same soup with the default routing: most of it is mid-size and overflows the warp-per-triangle queue instead
    _, got, _ = assert_parity(sc, stages=False)
    assert got["stats"]["mid_tris"] > 70000 and got["stats"]["regrow_count"] >= 1


def test_clip_record_overflow_is_repaired():
    # more fan triangles than the initial record capacity (65,536): the clipper must not leave keys that point at
    # records it could not write (the resolve would read past the array); the frame is re-run with larger queues
    rng = np.random.default_rng(9)
    n = 260000
    sc = scenes.config3(width=320, height=200, num_tris=n)
    v = sc.vertices.copy()
    v[:, 0:3] = (rng.random((n * 3, 3)) - 0.5) * np.array([6.0, 6.0, 1.2]) + np.array([0.0, 0.0, 0.3])
    sc["vertices"] = v
    _, got, _ = assert_parity(sc, stages=False)
    assert got["stats"]["clip_records"] > 65536 and got["stats"]["regrow_count"] >= 1, got["stats"]


def test_error_behaviour():
    from edxraster_b200 import renderer as R
    from edxraster_b200._lib import EdxError, EDX_ERR_UNSUPPORTED, EDX_ERR_INVALID
    r = R.Renderer(0)
    with pytest.raises(EdxError) as e:
        r.RenderMesh(type("M", (), {"_h": None})())
    assert e.value.code == EDX_ERR_INVALID
    r.Initialize(64, 64)
    with pytest.raises(EdxError) as e:
        r.SetMSAAMode(6)                           # tables end at 32x (FrameBuffer.cpp:107-191)
    assert e.value.code == EDX_ERR_INVALID
    r.SetMSAAMode(0)
    with pytest.raises(EdxError) as e:
        r.Initialize(8192, 8192)                   # 28.4 edge functions would overflow int32
    assert e.value.code == EDX_ERR_UNSUPPORTED
    with pytest.raises(EdxError):
        r.SetPixelShader(9)
    with pytest.raises(EdxError):
        r.SetOption("nonsense", 1)
    r.close()


@pytest.mark.parametrize("parts,msaa", [(2, 0), (3, 0), (2, 2)])
def test_sort_first_split_composites_to_the_full_frame(parts, msaa):
    # SURVEY.md §8e optional mode: each context owns the bins b % parts == part; together they are the frame
    import torch
    from edxraster_b200 import farm, renderer as R
    sc = scenes.config4(width=640, height=360, quads_x=240, quads_z=192)
    ref = parity.render_oracle(sc, msaa=msaa)
    colors, depths = [], []
    for part in range(parts):
        r = R.Renderer(0)
        r.Initialize(sc.width, sc.height)
        r.SetMSAAMode(msaa)
        r.SetTransform(sc.mv, sc.proj, sc.raster)
        r.SetPixelShader(sc.shader)
        r.SetScreenPartition(part, parts)
        m = r.CreateMesh(sc.vertices, sc.indices)
        for _ in range(2):                        # twice: the untouched foreign bins must stay untouched
            r.RenderMesh(m)
        colors.append(torch.from_numpy(r.GetBackBuffer().copy()))
        depths.append(torch.from_numpy(r.GetDepthBuffer()))
        m.Release()
        r.close()
    color = farm.composite_sort_first(torch.stack(colors)).numpy()
    depth = farm.composite_sort_first(torch.stack(depths)).numpy()
    np.testing.assert_array_equal(depth.view(np.uint32), ref["depth"].view(np.uint32))
    assert np.abs(color.astype(np.int32) - ref["color"].astype(np.int32)).max() <= 1
    # a partition really leaves foreign bins alone (cleared colour there)
    foreign = ~farm.bin_owner_mask(sc.width, sc.height, 0, parts).numpy()
    assert (colors[0].numpy()[foreign] == 0).all()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_tile_path_stress_many_candidates_per_bin(seed):
    # thousands of large overlapping triangles per 64x64 bin: exercises the candidate / survivor lists filling
    # up and flushing several times per bin (a barrier-divergence race hid here once: the loop decisions must be
    # taken from a count read before a barrier)
    rng = np.random.default_rng(seed)
    n = 60000
    c = rng.random((n, 1, 2)) * np.array([512, 384])
    p = c + (rng.random((n, 3, 2)) - 0.5) * rng.choice([40.0, 120.0, 300.0], size=(n, 1, 1))
    a, b = p[:, 0] - p[:, 2], p[:, 1] - p[:, 2]
    flip = (a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]) < 0
    p[flip, 0], p[flip, 1] = p[flip, 1].copy(), p[flip, 0].copy()
    sc = raster_scene(p, 0.05 + 0.9 * rng.random((n, 3)), 512, 384)
    ref = parity.render_oracle(sc)
    for opts in ({"mid_max": 0}, {"mid_max": 0, "hiz": 0}, {"mid_max": 0, "small_max": 4}, {}, {"bin_min": 0}, {"bin_min": -1, "mid_max": 0},
                 {"bin_min": 100, "mid_max": 0, "hiz": 0}):
        got = parity.render_gpu(sc, options=opts, stages=False)
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), (opts, rep)
        assert got["stats"]["binned_tris"] > (20000 if "mid_max" in opts else 8000)
        # per-bin lists (stage a7) are built for long tile-path lists, and only for them
        if opts.get("bin_min", 1) <= 0:
            assert (got["stats"]["bin_pairs"] > got["stats"]["binned_tris"]) == (opts["bin_min"] < 0), (opts, got["stats"])


def _fuzz_scene(seed):
    """Random soup: mixed sizes from sub-pixel to several screens, random depths incl. behind the camera,
    random perspective camera, so every clip plane, the w <= 0 drop, both raster paths and HiZ get hit."""
    from edxraster_b200 import camera as cam
    rng = np.random.default_rng(seed)
    w, h = int(rng.choice([160, 256, 322])), int(rng.choice([96, 192, 130]))
    n = 2500
    centre = (rng.random((n, 1, 3)) - 0.5) * np.array([6.0, 4.0, 10.0]) + np.array([0.0, 0.0, 3.0])
    size = rng.choice([0.01, 0.05, 0.3, 2.0, 12.0], size=(n, 1, 1), p=[0.3, 0.3, 0.2, 0.15, 0.05])
    pos = centre + (rng.random((n, 3, 3)) - 0.5) * size
    v = np.zeros((n * 3, 8), np.float32)
    v[:, 0:3] = pos.reshape(-1, 3)
    v[:, 3:6] = rng.normal(size=(n * 3, 3))
    v[:, 6:8] = rng.random((n * 3, 2))
    eye = (rng.random(3) - 0.5) * np.array([2.0, 2.0, 2.0]) + np.array([0.0, 0.0, -2.0])
    c = cam.Camera(eye, (0.0, 0.0, 3.0), (0.0, 1.0, 0.0), w, h, float(rng.choice([40.0, 65.0, 100.0])),
                   float(rng.choice([0.05, 0.5, 2.0])), float(rng.choice([6.0, 50.0])))
    sc = scenes.Scene(name="fuzz%d" % seed, width=w, height=h, vertices=v,
                      indices=np.arange(n * 3, dtype=np.uint32).reshape(-1, 3), mv=c.view, proj=c.proj, raster=c.raster,
                      shader=int(rng.choice([0, 1, 2, 3])))
    msaa = int(rng.choice([0, 0, 1, 2, 3, 4, 5]))
    sc["front_end"] = int(rng.choice([0, 1, 2]))
    sc["mid_max"] = int(rng.choice([0, 24, 64, 64, 300]))
    if sc.shader == 3 and rng.random() < 0.8:            # textured: random slots, sizes, filter, texcoord range
        v[:, 6:8] = (v[:, 6:8] - 0.5) * float(rng.choice([1.0, 4.0, 40.0]))
        tex = []
        for k in range(int(rng.integers(1, 4))):
            if rng.random() < 0.25:
                tex.append(("constant", tuple(rng.random(3))))
            else:
                tex.append(("image", scenes.noise_texture(int(rng.choice([1, 2, 5, 16, 33, 128])), int(rng.choice([1, 3, 8, 64])), seed + k, cell=2)))
        sc["textures"], sc["tex_ids"] = tex, rng.integers(0, len(tex), n).astype(np.uint32)
        sc["tex_filter"] = int(rng.integers(0, 6))
    return sc, msaa


@pytest.mark.parametrize("seed", list(range(100, 124)))
def test_fuzz_random_scenes(seed):
    sc, msaa = _fuzz_scene(seed)
    ref = parity.render_oracle(sc, msaa=msaa)
    got = parity.render_gpu(sc, msaa=msaa, stages=(msaa == 0), options={"front_end": sc["front_end"], "cluster_cull": 2 * (seed & 1), "mid_max": sc["mid_max"]})
    rep = parity.compare(ref, got)
    assert parity.is_parity(rep), (seed, msaa, rep)


# ---- several frames in flight (FrameRing: contexts on their own streams sharing one mesh) ----

def test_frame_ring_frames_in_flight_match_the_oracle():
    import copy
    from edxraster_b200 import renderer as R
    base = scenes.by_name("C4", 0.02)                   # ~100k-triangle terrain, clipped at the screen edges
    views = scenes.config5_views(base, 7)
    ring = R.FrameRing(0, depth=3)
    ring.Initialize(base.width, base.height)
    ring.SetPixelShader(base.shader)
    mesh = ring.CreateMesh(base.vertices, base.indices)
    for lane in ring.lanes:
        lane.SetCaptureIds(True)
    refs = []
    for v in views:
        sc = copy.copy(base)
        sc.mv, sc.proj, sc.raster = v
        refs.append(parity.render_oracle(sc))
    # keep three frames in flight: read frame i - 2 back after submitting frame i
    got = {}
    tickets = []
    for i, v in enumerate(views):
        # both spellings of the transform: three matrices, or marshalled once (PackedTransform)
        tickets.append(ring.Submit(mesh, R.PackedTransform(*v)) if i % 2 else ring.Submit(mesh, *v))
        if i >= 2:
            t = tickets[i - 2]
            lane = ring.lanes[t % 3]
            got[t] = {"color": ring.GetBackBuffer(t).copy(), "depth": ring.GetDepthBuffer(t), "winner": lane.GetWinnerIds(), "derived": lane.DerivedState()}
    for t in tickets[-2:]:
        lane = ring.lanes[t % 3]
        got[t] = {"color": ring.GetBackBuffer(t).copy(), "depth": ring.GetDepthBuffer(t), "winner": lane.GetWinnerIds(), "derived": lane.DerivedState()}
    with pytest.raises(ValueError):
        ring.GetBackBuffer(tickets[0])                  # left the ring long ago
    for t, ref in zip(tickets, refs):
        rep = parity.compare(ref, got[t])
        assert parity.is_parity(rep), (t, rep)
    assert len({g["depth"].tobytes() for g in got.values()}) == len(views)     # the views really differ
    mesh.Release()
    ring.close()


def test_overflow_in_an_unsynchronised_earlier_frame_is_reported():
    # frame A overflows the tile-path queue, frame B (submitted right after, no synchronising call in between)
    # does not: the next synchronising call must say that a frame was lost, once, and rendering A again works
    from edxraster_b200 import renderer as R
    from edxraster_b200._lib import EdxError, EDX_ERR_OVERFLOW
    rng = np.random.default_rng(5)
    n = 120000
    c = rng.random((n, 1, 2)) * np.array([1280, 720])
    p = c + (rng.random((n, 3, 2)) - 0.5) * 90
    a, b = p[:, 0] - p[:, 2], p[:, 1] - p[:, 2]
    flip = (a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]) < 0
    p[flip, 0], p[flip, 1] = p[flip, 1].copy(), p[flip, 0].copy()
    big = raster_scene(p, 0.1 + 0.8 * rng.random((n, 3)), 1280, 720)
    small = raster_scene(p[:50], 0.1 + 0.8 * rng.random((50, 3)), 1280, 720)
    r = R.Renderer(0)
    r.Initialize(1280, 720)
    r.SetOption("mid_max", 0)                           # everything above the direct path's box goes to the tile-path queue
    r.SetPixelShader(big.shader)
    r.SetTransform(big.mv, big.proj, big.raster)
    mb, msm = r.CreateMesh(big.vertices, big.indices), r.CreateMesh(small.vertices, small.indices)
    r.RenderMesh(mb)
    r.RenderMesh(msm)
    with pytest.raises(EdxError) as e:
        r.Synchronize()
    assert e.value.code == EDX_ERR_OVERFLOW and "1 frame" in str(e.value)
    r.Synchronize()                                     # reported once
    r.RenderMesh(mb)                                    # the queues were grown: this time the frame is complete
    got = r.GetDepthBuffer()
    assert r.GetStats()["regrow_count"] == 1
    ref = parity.render_oracle(big)
    assert (ref["depth"].view(np.uint32) != got.view(np.uint32)).sum() == 0
    r.close()


# ---- textured LambertianAlbedoPixelShader + filter modes (Core/Shader.h:209-244) ----

@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4, 5])
def test_textured_plane_every_filter(filt):
    # grazing-angle plane: every mip level, strongly anisotropic footprints, near-plane clipping of the front row
    sc = scenes.textured_plane(tex_filter=filt)
    _, got, rep = assert_parity(sc)
    assert got["stats"]["clipped_tris"] > 0
    assert len(np.unique(got["color"].reshape(-1, 4), axis=0)) > 5000


@pytest.mark.parametrize("filt", [0, 2, 4])
def test_textured_sphere_slots_and_odd_sizes(filt):
    # constant + 37x21 + 2x2 textures assigned per triangle, texcoords beyond [0, 1) (repeat)
    assert_parity(scenes.textured_sphere(tex_filter=filt), stages=False)


def test_textured_terrain_with_clipping_and_msaa():
    sc = scenes.config4(width=640, height=360, quads_x=200, quads_z=160)
    sc["shader"] = scenes.SHADER_LAMBERT_ALBEDO
    sc["textures"] = [("image", scenes.noise_texture(64, 64, 9)), ("constant", (0.2, 0.8, 0.4)), ("image", scenes.noise_texture(16, 128, 10))]
    sc["tex_ids"] = (np.arange(sc.num_tris, dtype=np.uint32) // 3) % 3
    v = sc.vertices.copy()
    v[:, 6:8] = v[:, [0, 2]] * 0.7                     # world-space planar mapping
    sc["vertices"] = v
    for filt, msaa in ((2, 0), (3, 0), (2, 2)):
        sc["tex_filter"] = filt
        assert_parity(sc, stages=False, msaa=msaa)


def test_device_mip_chain_equals_the_oracle_chain():
    from edxraster_b200 import renderer as R
    from oracle import orc
    imgs = [scenes.noise_texture(128, 64, 1), scenes.noise_texture(37, 21, 2, cell=3), scenes.noise_texture(1, 5, 3, cell=1)]
    o = orc.Oracle(16, 16, 1)
    o.set_textures([("image", i) for i in imgs])
    r = R.Renderer(0)
    r.Initialize(16, 16)
    sc = scenes.config1(width=16, height=16, slices=4, stacks=4)
    m = r.CreateMesh(sc.vertices, sc.indices)
    m.SetTextures([("image", i) for i in imgs])
    for slot in range(3):
        ref = o.tex_mips(slot)
        for level, want in enumerate(ref):
            np.testing.assert_array_equal(m.TextureLevel(slot, level), want)
    m.Release()
    r.close()


def test_texture_api_errors_and_replacement():
    from edxraster_b200 import renderer as R
    from edxraster_b200._lib import EdxError, EDX_ERR_INVALID
    sc = scenes.textured_sphere()
    r = R.Renderer(0)
    r.Initialize(sc.width, sc.height)
    m = r.CreateMesh(sc.vertices, sc.indices)
    with pytest.raises(EdxError) as e:
        m.SetTextures([("constant", (1, 1, 1))], np.full(sc.num_tris, 1, np.uint32))     # slot 1 does not exist
    assert e.value.code == EDX_ERR_INVALID
    with pytest.raises(ValueError):
        m.SetTextures([("image", np.zeros((4, 4, 3), np.uint8))])
    with pytest.raises(EdxError):
        m.TextureLevel(0, 0)
    # replacing and removing textures: the last state wins, none = the context's constant albedo
    m.SetTextures(sc["textures"], sc["tex_ids"])
    m.SetTextures([])
    r.SetTransform(sc.mv, sc.proj, sc.raster)
    r.SetPixelShader(scenes.SHADER_LAMBERT_ALBEDO)
    r.RenderMesh(m)
    got = r.GetBackBuffer().copy()
    plain = dict(sc)
    plain.pop("textures"); plain.pop("tex_ids")
    ref = parity.render_oracle(scenes.Scene(plain))
    assert np.abs(got.astype(np.int32) - ref["color"].astype(np.int32)).max() <= 1
    m.Release()
    r.close()


def test_farm_render_views_through_a_frame_ring_into_torch_tensors():
    # the frame farm's per-rank loop (farm.render_views) with three views in flight, rendering straight into the
    # torch tensors a gather would send
    import copy
    import torch
    from edxraster_b200 import farm, renderer as R
    base = scenes.by_name("C4", 0.01)
    views = scenes.config5_views(base, 5)
    ring = R.FrameRing(0, depth=3)
    ring.Initialize(base.width, base.height)
    ring.SetPixelShader(base.shader)
    mesh = ring.CreateMesh(base.vertices, base.indices)
    dev = torch.device("cuda", 0)
    out = torch.zeros((len(views), base.height, base.width, 4), dtype=torch.uint8, device=dev)
    farm.render_views(ring, mesh, [R.PackedTransform(*v) for v in views], out, shaded=True)
    got = out.cpu().numpy()
    single = R.Renderer(0)
    single.Initialize(base.width, base.height)
    single.SetPixelShader(base.shader)
    out1 = torch.zeros_like(out)
    farm.render_views(single, mesh, views, out1, shaded=True)          # the shared mesh, one frame at a time, plain matrices
    np.testing.assert_array_equal(got, out1.cpu().numpy())
    for k, v in enumerate(views):
        sc = copy.copy(base)
        sc.mv, sc.proj, sc.raster = v
        ref = parity.render_oracle(sc)
        assert np.abs(got[k].astype(np.int32) - ref["color"].astype(np.int32)).max() <= 1, k
    assert len({g.tobytes() for g in got}) == len(views)
    mesh.Release()
    single.close()
    ring.close()


def test_default_state_is_the_reference_default():
    # Renderer::Initialize installs LambertianAlbedoPixelShader (Renderer.cpp:41) and TriLinear (RenderStates.h:60);
    # a mesh without textures is shaded with the constant 0.9 white of LoadSphere / LoadPlane (Mesh.cpp:47,66)
    from edxraster_b200 import renderer as R
    sc = scenes.config1(width=320, height=180, slices=24, stacks=24)
    r = R.Renderer(0)
    r.Initialize(sc.width, sc.height)
    r.SetTransform(sc.mv, sc.proj, sc.raster)
    m = r.CreateMesh(sc.vertices, sc.indices)
    r.RenderMesh(m)
    got = r.GetBackBuffer().copy()
    ref = parity.render_oracle(sc, shader=scenes.SHADER_LAMBERT_ALBEDO)
    assert np.abs(got.astype(np.int32) - ref["color"].astype(np.int32)).max() <= 1
    assert (got[..., 3] == 255).sum() > 3000
    m.Release()
    r.close()


# ---- round 2 additions ----

@pytest.mark.parametrize("level", [1, 3, 5])
def test_msaa_on_an_odd_sized_frame(level):
    # W, H not multiples of 2 (nor of the 16 / 64 pixel tiles): partial quads, tiles and bins at both far edges
    sc = scenes.config4(width=327, height=201, quads_x=120, quads_z=96)
    assert_parity(sc, msaa=level, stages=False)
    assert_parity(scenes.config1(width=327, height=201, slices=40, stacks=40), msaa=level, stages=False)


def test_depth_before_the_first_frame_is_the_clear_value():
    from edxraster_b200 import renderer as R
    r = R.Renderer(0)
    r.Initialize(96, 64)
    assert (r.GetDepthBuffer() == 1.0).all()            # FrameBuffer::Init / Clear, FrameBuffer.cpp:103
    r.SetCaptureIds(True)
    r.SetMSAAMode(2)
    d, _ = r.GetSample(3)
    assert (d == 1.0).all()
    r.close()


def test_overflowed_frame_is_rerun_as_submitted():
    # the frame overflows the mid-size queue; before the synchronising call the caller already sets the NEXT frame's
    # transform and shader. The repair must render the frame that was submitted, not the context's current state.
    from edxraster_b200 import renderer as R
    rng = np.random.default_rng(5)
    n = 120000
    c = rng.random((n, 1, 2)) * np.array([1280, 720])
    p = c + (rng.random((n, 3, 2)) - 0.5) * 90
    a, b = p[:, 0] - p[:, 2], p[:, 1] - p[:, 2]
    flip = (a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]) < 0
    p[flip, 0], p[flip, 1] = p[flip, 1].copy(), p[flip, 0].copy()
    sc = raster_scene(p, 0.1 + 0.8 * rng.random((n, 3)), 1280, 720)
    sc["shader"] = 1
    other = scenes.config1(width=1280, height=720)
    r = R.Renderer(0)
    r.Initialize(1280, 720)
    r.SetCaptureIds(True)
    r.SetPixelShader(1)
    r.SetTransform(sc.mv, sc.proj, sc.raster)
    m = r.CreateMesh(sc.vertices, sc.indices)
    r.RenderMesh(m)
    r.SetTransform(other.mv, other.proj, other.raster)      # state of a frame that is not submitted yet
    r.SetPixelShader(0)
    got = {"color": r.GetBackBuffer().copy(), "depth": r.GetDepthBuffer(), "winner": r.GetWinnerIds()}
    assert r.GetStats()["regrow_count"] >= 1
    ref = parity.render_oracle(sc)
    got["derived"] = ref["derived"]
    rep = parity.compare(ref, got)
    assert parity.is_parity(rep), rep
    r.close()


def test_streamed_meshes_on_three_contexts_match_the_oracle():
    # the end-to-end path bench.py times (edx_mesh_update from pinned host memory + edx_set_transform + edx_render_mesh +
    # read-back, three contexts in flight on their own streams), checked instead of timed: every frame of every lane
    import torch
    from edxraster_b200 import renderer as R
    frames = [scenes.config2(width=640, height=360, num_tris=30000, seed=s) for s in (1, 2, 3, 4, 5, 6, 7)]
    nv, nt = frames[0].num_verts, frames[0].num_tris
    lanes = []
    for k in range(3):
        st = torch.cuda.Stream()
        r = R.Renderer(0)
        r.SetStream(st.cuda_stream)
        r.Initialize(640, 360)
        r.SetPixelShader(0)
        r.SetCaptureIds(True)
        lanes.append((st, r, r.CreateMesh(frames[0].vertices, frames[0].indices)))
    hv = [torch.from_numpy(np.ascontiguousarray(f.vertices)).pin_memory() for f in frames]
    hi = [torch.from_numpy(np.ascontiguousarray(f.indices).view(np.int32)).pin_memory() for f in frames]
    out = [torch.empty((360, 640), dtype=torch.float32).pin_memory() for _ in frames]
    K = 3
    for i in range(len(frames) + K - 1):
        if i < len(frames):
            _, r, mesh = lanes[i % K]
            mesh.update(hv[i].data_ptr(), nv, hi[i].data_ptr(), nt)
            r.SetTransform(frames[i].mv, frames[i].proj, frames[i].raster)
            r.RenderMesh(mesh)
        j = i - K + 1
        if j >= 0:
            lanes[j % K][1].ReadDepthInto(out[j].data_ptr())
    for f, o in zip(frames, out):
        ref = parity.render_oracle(f)
        assert (ref["depth"].view(np.uint32) != o.numpy().view(np.uint32)).sum() == 0
    assert len({o.numpy().tobytes() for o in out}) == len(frames)
    for _, r, mesh in lanes:
        mesh.Release()
        r.close()


def test_m1_mid_size_stress_every_routing():
    # M1 (bench.py --workload M1, reduced): triangles of 32..128 px. Default routing sends most to the warp-per-triangle
    # path; mid_max = 0 is round 1's routing (everything on the tile path, every bin sweeps the whole list).
    sc = scenes.stress_m1(width=1280, height=720, num_tris=60000)
    ref = parity.render_oracle(sc)
    for opts in ({}, {"mid_max": 0, "small_max": 32}, {"mid_max": 128}, {"sort_big": 0}, {"bin_min": 0}, {"bin_min": 0, "mid_max": 0}, {"bin_min": 1000, "mid_max": 0}):
        got = parity.render_gpu(sc, options=opts, stages=False)
        rep = parity.compare(ref, got)
        assert parity.is_parity(rep), (opts, rep)
        if "bin_min" in opts:           # long tile-path lists are binned (stage a7) unless switched off
            assert (got["stats"]["bin_pairs"] > 0) == (opts["bin_min"] > 0 and got["stats"]["binned_tris"] >= opts["bin_min"]), (opts, got["stats"])
    assert got["stats"]["mid_tris"] + got["stats"]["binned_tris"] > 50000


@pytest.mark.parametrize("skip_tile", [1, 0])
def test_idle_kernels_are_left_out_safely(skip_tile):
    # mid_kernel / sort_big_kernel / tile_kernel are left out of a frame when the previous frame of the same mesh had no
    # work for them (skip_idle, skip_tile). Same mesh, five views: far away twice (every triangle tiny), then close up
    # three times: mid-size and large triangles appear in a frame that lacks the kernel meant for them and must take the
    # other path (without mid_kernel: the tile path; without tile_kernel: mid_kernel, at any size). All five exact.
    import copy
    from edxraster_b200 import camera as cam, renderer as R
    base = scenes.config1(width=640, height=360, slices=80, stacks=80)
    r = R.Renderer(0)
    r.Initialize(base.width, base.height)
    r.SetCaptureIds(True)
    r.SetPixelShader(1)
    r.SetOption("skip_tile", skip_tile)
    m = r.CreateMesh(base.vertices, base.indices)
    launches, stats = [], []
    for eye_z in (-60.0, -60.0, -1.6, -1.6, -1.6):
        c = cam.Camera((0.0, 0.0, eye_z), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), base.width, base.height, 65.0, 0.01, 100.0)
        sc = copy.copy(base)
        sc.mv, sc.proj, sc.raster = c.view, c.proj, c.raster
        r.SetTransform(sc.mv, sc.proj, sc.raster)
        r.RenderMesh(m)
        got = {"color": r.GetBackBuffer().copy(), "depth": r.GetDepthBuffer(), "winner": r.GetWinnerIds(), "derived": r.DerivedState()}
        launches.append(r.LastLaunchList())
        stats.append(r.GetStats())
        rep = parity.compare(parity.render_oracle(sc), got)
        assert parity.is_parity(rep), (eye_z, rep)
    if skip_tile:
        assert "tile_kernel" in launches[0] and "mid_kernel" in launches[0]
        # frame 0 put nothing on the tile path: frames 1 and 2 have no tile_kernel (a shaded frame then needs no resolve kernel
        # either: shade_kernel reads the keys itself), with mid_kernel as the catch-all
        for i in (1, 2):
            assert "tile_kernel" not in launches[i] and "lean_resolve_kernel" not in launches[i], launches[i]
            assert "shade_kernel" in launches[i] and "mid_kernel" in launches[i], launches[i]
        # frame 2 (close up) had large triangles: the tile kernel is back, and they are on its path
        assert "tile_kernel" in launches[3] and "tile_kernel" in launches[4] and stats[3]["binned_tris"] > 0
    else:
        assert "mid_kernel" in launches[0] and "mid_kernel" not in launches[1]       # frame 0 had no mid-size triangle
        assert "mid_kernel" not in launches[2] and "mid_kernel" in launches[3]       # frame 2 diverted some: back in frame 3
    r.close()


def test_frame_sink_pushes_every_frame_and_counts_them():
    # edx_set_frame_sink / edx_set_frame_sink_signal on one GPU (the store is the context's own memory; over NVLink it is a
    # peer's): three views, each pushed into its own slot by the copy engine behind the frame; the signal word counts them
    import copy
    from edxraster_b200 import camera as cam, renderer as R
    base = scenes.config1(width=320, height=200, slices=40, stacks=40)
    r = R.Renderer(0)
    r.Initialize(base.width, base.height)
    r.SetPixelShader(1)
    m = r.CreateMesh(base.vertices, base.indices)
    fb = base.width * base.height * 4
    store = r.DeviceAlloc(3 * 2 * fb + 16)
    word = store + 3 * 2 * fb
    r.SetFrameSinkSignal(word)
    scs = []
    for i, eye_z in enumerate((-3.0, -2.2, -1.7)):
        c = cam.Camera((0.3 * i, 0.0, eye_z), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), base.width, base.height, 65.0, 0.01, 100.0)
        sc = copy.copy(base)
        sc.mv, sc.proj, sc.raster = c.view, c.proj, c.raster
        scs.append(sc)
        r.SetTransform(sc.mv, sc.proj, sc.raster)
        r.SetFrameSink(store + (2 * i) * fb, store + (2 * i + 1) * fb)
        r.RenderMesh(m)                   # not synchronised: the pushes are stream-ordered behind each frame
    r.Synchronize()
    assert int(r.ReadDevice(word, 4).view(np.uint32)[0]) == 3
    for i, sc in enumerate(scs):
        ref = parity.render_oracle(sc)
        color = r.ReadDevice(store + (2 * i) * fb, fb).reshape(base.height, base.width, 4)
        depth = r.ReadDevice(store + (2 * i + 1) * fb, fb).view(np.float32).reshape(base.height, base.width)
        assert np.array_equal(depth.view(np.uint32), np.asarray(ref["depth"], dtype=np.float32).reshape(base.height, base.width).view(np.uint32)), i
        assert np.abs(color.astype(int) - np.asarray(ref["color"]).reshape(base.height, base.width, 4).astype(int)).max() <= 1, i
    r.SetFrameSink(0, 0)
    r.SetFrameSinkSignal(0)
    r.DeviceFree(store)
    r.close()
