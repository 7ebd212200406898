"""Stage times of one config under different tuning knobs (GPU)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C3")
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
sc = scenes.by_name(a.workload, a.scale)
r = R.Renderer(0)
r.Initialize(sc.width, sc.height)
r.SetTransform(sc.mv, sc.proj, sc.raster)
r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
r.SetProfiling(True)
for opts in ({"cluster_cull": 1}, {"cluster_cull": 0}):
    for k, v in opts.items():
        r.SetOption(k, v)
    acc = {"geom": 0, "clip": 0, "tile": 0, "total": 0}
    for i in range(6):
        r.RenderMesh(m); r.Synchronize()
        if i >= 2:
            st = r.GetStats()
            for k in acc: acc[k] += st["stage_ms"][k] / 4
    r.SetProfiling(False)
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(50): r.RenderMesh(m)
    ms = r.TimerEnd() / 50
    r.SetProfiling(True)
    print(a.workload, opts, {k: round(v * 1000, 1) for k, v in acc.items()}, "binned", st["binned_tris"], "back-to-back us/frame %.1f" % (ms * 1000), flush=True)
