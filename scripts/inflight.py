"""Throughput with k frames in flight on one GPU (k contexts, each on its own stream, alternating frames)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from edxraster_b200 import renderer as R, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = scenes.by_name(name, 1.0)
nt = len(sc.indices)
torch.cuda.set_device(0)
ks = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3, 4]
warm = torch.empty(1 << 28, device='cuda')
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.5: warm.add_(1.0)
torch.cuda.synchronize()
for k in ks:
    ctx = []
    for i in range(k):
        r = R.Renderer(0); r.Initialize(sc.width, sc.height)
        r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
        ms = [r.CreateMesh(sc.vertices, sc.indices) for _ in range(2)]
        ctx.append((r, ms))
    frames = 400 if nt <= 2_000_000 else 60
    best = 1e9
    for rep in range(4):
        for r, _ in ctx: r.Synchronize()
        t = time.perf_counter()
        for f in range(frames):
            r, ms = ctx[f % k]
            r.RenderMesh(ms[(f // k) % 2])
        for r, _ in ctx: r.Synchronize()
        best = min(best, (time.perf_counter() - t) / frames * 1e6)
    print(f"{name} in-flight {k}: {best:.1f} us/frame  {nt/best:.0f} Mtris/s", flush=True)
    del ctx
