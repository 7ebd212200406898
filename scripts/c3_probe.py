"""Why are back-to-back C3 frames 2x slower in a process that has not initialised torch? Probe variants."""
import os, sys, time
sys.path.insert(0, ".")
mode = sys.argv[1]
if mode == "torch":
    import torch; torch.zeros(1, device="cuda")
if mode == "cudart_malloc":
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p(); rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(2 << 20))
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
if mode == "nopdl": r.SetOption("pdl", 0)
m = r.CreateMesh(sc.vertices, sc.indices)
for _ in range(5): r.RenderMesh(m)
r.Synchronize()
if mode == "sync_each":
    r.SetProfiling(True)
    tot = 0
    for _ in range(30):
        r.RenderMesh(m); r.Synchronize(); tot += r.GetStats()["stage_ms"]["total"]
    print(mode, f"{tot/30*1e3:.1f} us/frame (per-frame events, synchronised)")
else:
    r.TimerBegin()
    for _ in range(30): r.RenderMesh(m)
    print(mode, f"{r.TimerEnd()/30*1e3:.1f} us/frame")
