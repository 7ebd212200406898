import sys, os
sys.path.insert(0, ".")
mode = sys.argv[1]
env0 = dict(os.environ)
if mode != "none":
    import torch
    if mode == "init": torch.cuda.init()
    if mode == "tensor": torch.zeros(1, device="cuda")
    if mode == "after": pass
for k in os.environ:
    if os.environ.get(k) != env0.get(k): print("env changed:", k, os.environ[k])
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
def t():
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(30): r.RenderMesh(m)
    return r.TimerEnd() / 30 * 1e3
print(mode, f"{t():.1f} us/frame")
if mode == "after":
    torch.zeros(1, device="cuda"); print("after torch tensor:", f"{t():.1f} us/frame")
    r2 = R.Renderer(0); r2.Initialize(sc.width, sc.height); r2.SetTransform(sc.mv, sc.proj, sc.raster); r2.SetPixelShader(sc.shader)
    m2 = r2.CreateMesh(sc.vertices, sc.indices); r = r2; m = m2
    print("new context after torch tensor:", f"{t():.1f} us/frame")
