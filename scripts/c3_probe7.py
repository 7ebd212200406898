import sys
sys.path.insert(0, ".")
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
def t(tag):
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(30): r.RenderMesh(m)
    print(tag, f"{r.TimerEnd() / 30 * 1e3:.1f} us/frame", flush=True)
    r.RenderMesh(m); r.Synchronize()
t("before")
import torch
x = torch.zeros(1, device="cuda"); torch.cuda.synchronize()
t("after torch kernel")
