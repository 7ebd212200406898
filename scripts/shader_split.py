"""Frame time of a config with shading on vs depth-only (isolates raster from shade cost)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="C3"); a = ap.parse_args()
sc = scenes.by_name(a.workload, 1.0)
r = R.Renderer(0); r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster)
m = r.CreateMesh(sc.vertices, sc.indices)
for shader in (sc.shader, 0, 2):
    r.SetPixelShader(shader)
    for _ in range(3): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(10): r.RenderMesh(m)
    print(a.workload, "shader", shader, "ms/frame %.4f" % (r.TimerEnd() / 10), flush=True)
