"""Micro-benchmark of the tile path: one 64x64 bin, N screen-covering triangles, HiZ off -> every warp
rasterises every triangle into its tile; reports cycles-equivalent per (triangle, tile) pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes
r = R.Renderer(0)
for (w, h, n, hiz) in ((64, 64, 250, 0), (64, 64, 1000, 0), (64, 64, 1000, 1), (128, 128, 1000, 0), (512, 512, 1000, 0)):
    sc = scenes.config3(width=w, height=h, num_tris=n)
    r.Initialize(w, h); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(0); r.SetOption("hiz", hiz)
    m = r.CreateMesh(sc.vertices, sc.indices)
    r.SetProfiling(True)
    t = 0.0
    for i in range(6):
        r.RenderMesh(m); r.Synchronize()
        if i >= 2: t += r.GetStats()["stage_ms"]["tile"] / 4
    st = r.GetStats()
    pairs_per_warp = st["binned_tris"]
    print(w, h, "tris", n, "hiz", hiz, "binned", st["binned_tris"], "tile ms %.4f" % t, "-> us per (tri,tile) per warp: %.3f" % (t * 1000 / max(1, pairs_per_warp)), flush=True)
    r.SetProfiling(False); m.Release()
