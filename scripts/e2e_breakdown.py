"""Where the end-to-end step (host mesh -> device, render, frame -> host) spends its time. GPU box only."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from edxraster_b200 import renderer as R, scenes

def wall(fn, n=20):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

sc = scenes.by_name(sys.argv[1] if len(sys.argv) > 1 else "C2", 1.0)
nv, nt = len(sc.vertices), len(sc.indices)
torch.cuda.set_device(0)
r = R.Renderer(0); r.Initialize(sc.width, sc.height)
r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(R.SHADER_BLINN_PHONG)
mesh = r.CreateMesh(sc.vertices, sc.indices)
hv = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
hi = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
dv = torch.empty_like(hv, device="cuda"); di = torch.empty_like(hi, device="cuda")
dc = torch.empty((sc.height, sc.width, 4), dtype=torch.uint8, device="cuda"); hc = torch.empty_like(dc, device="cpu").pin_memory()
print(f"nv {nv} nt {nt} h2d {nv*32+nt*12} B d2h {dc.numel()} B")
t = wall(lambda: dv.copy_(hv, non_blocking=True)); print(f"raw H2D vertices  {t:.3f} ms  {nv*32/t/1e6:.1f} GB/s")
t = wall(lambda: di.copy_(hi, non_blocking=True)); print(f"raw H2D indices   {t:.3f} ms  {nt*12/t/1e6:.1f} GB/s")
t = wall(lambda: hc.copy_(dc, non_blocking=True)); print(f"raw D2H frame     {t:.3f} ms  {dc.numel()/t/1e6:.1f} GB/s")
def both():
    dv.copy_(hv, non_blocking=True); di.copy_(hi, non_blocking=True); hc.copy_(dc, non_blocking=True)
t = wall(both); print(f"raw all three serial {t:.3f} ms")
t = wall(lambda: (mesh.update(hv.data_ptr(), nv, hi.data_ptr(), nt), r.Synchronize())); print(f"mesh.update+sync  {t:.3f} ms")
t = wall(lambda: (r.RenderMesh(mesh), r.Synchronize())); print(f"render+sync       {t:.3f} ms")
t = wall(lambda: r.GetBackBuffer()); print(f"GetBackBuffer     {t:.3f} ms")
def step():
    mesh.update(hv.data_ptr(), nv, hi.data_ptr(), nt); r.SetTransform(sc.mv, sc.proj, sc.raster); r.RenderMesh(mesh); r.GetBackBuffer()
t = wall(step); print(f"full step         {t:.3f} ms  {nt/t/1e3:.0f} Mtris/s")
