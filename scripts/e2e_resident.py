"""Resident-mesh end-to-end step (transform in, render, frame to pinned host) with k lanes: where does time go?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C2", 1.0)
nt = len(sc.indices)
torch.cuda.set_device(0)
xf = R.PackedTransform(sc.mv, sc.proj, sc.raster)
warm = torch.empty(1 << 28, device="cuda")
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0: warm.add_(1.0)
torch.cuda.synchronize()
for K in (1, 2, 3):
    lanes = []
    for _ in range(K):
        r = R.Renderer(0); r.Initialize(sc.width, sc.height); r.SetPixelShader(sc.shader); r.SetTransform(xf); lanes.append(r)
    mesh = lanes[0].CreateMesh(sc.vertices, sc.indices)
    houts = [torch.empty((sc.height, sc.width), dtype=torch.float32).pin_memory() for _ in range(K)]
    def run(n):
        for i in range(n + K - 1):
            if i < n:
                lanes[i % K].SetTransform(xf); lanes[i % K].RenderMesh(mesh)
            if i >= K - 1:
                j = (i - K + 1) % K
                lanes[j].ReadDepthInto(houts[j].data_ptr())
    run(10); torch.cuda.synchronize()
    t = time.perf_counter(); run(100); torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 100 * 1e3
    # the pieces alone
    t = time.perf_counter()
    for i in range(100): lanes[0].ReadDepthInto(houts[0].data_ptr())
    rd = (time.perf_counter() - t) / 100 * 1e3
    print(f"K={K}: step {dt:.3f} ms  ({nt/dt/1e3:.0f} Mtris/s)   read-back alone {rd:.3f} ms", flush=True)
    mesh.Release()
    for r in lanes: r.close()
