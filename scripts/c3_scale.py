import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes
r = R.Renderer(0)
for (w, h, n) in ((3840, 2160, 250), (3840, 2160, 500), (3840, 2160, 1000), (3840, 2160, 2000), (3840, 2160, 4000), (1920, 1080, 2000), (960, 540, 2000)):
    sc = scenes.config3(width=w, height=h, num_tris=n)
    r.Initialize(w, h); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(0)
    m = r.CreateMesh(sc.vertices, sc.indices)
    for _ in range(3): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(5): r.RenderMesh(m)
    print(w, h, n, "ms/frame %.4f" % (r.TimerEnd() / 5), flush=True)
    m.Release()
