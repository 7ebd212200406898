"""Frame time vs MSAA level for reduced configs (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes
r = R.Renderer(0)
cases = {"C1": scenes.config1(), "C2": scenes.config2(), "C3_1080p": scenes.config3(width=1920, height=1080, num_tris=2000), "C4": scenes.config4()}
for name, sc in cases.items():
    r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
    m = r.CreateMesh(sc.vertices, sc.indices)
    for level in (0, 1, 2, 3):
        r.SetMSAAMode(level)
        for _ in range(3): r.RenderMesh(m)
        r.Synchronize(); r.TimerBegin()
        for _ in range(10): r.RenderMesh(m)
        print(name, "%2dx" % (1 << level), "ms/frame %.3f" % (r.TimerEnd() / 10), flush=True)
    r.SetMSAAMode(0)
    m.Release()
