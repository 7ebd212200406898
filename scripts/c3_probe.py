"""The carve-out interaction of DESIGN.md section 7: C3 frame time in a process that has / has not run a kernel of
another CUDA module, with the tile kernel's real CTA residency and the real SM clock next to it. GPU box only.

    python scripts/c3_probe.py none|import|tensor|after [clip_carveout]

none: no torch; import: torch imported, CUDA untouched; tensor: one torch kernel before the context is created;
after: measure, run one torch kernel, measure again with the SAME context and buffers. clip_carveout: 0 auto
(default), 1 = always prefer L1 (shows the slow state), 2 = always prefer shared memory.
EDX_DEBUG_PRINT=1 prints the residency / clock line."""
import os, sys
sys.path.insert(0, ".")
mode = sys.argv[1] if len(sys.argv) > 1 else "none"
carve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if mode != "none":
    import torch
    if mode == "tensor":
        torch.zeros(1, device="cuda")
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
r.SetOption("clip_carveout", carve)
m = r.CreateMesh(sc.vertices, sc.indices)

def t(tag):
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(30): r.RenderMesh(m)
    ms = r.TimerEnd() / 30
    print(f"{tag}: {ms * 1e3:.1f} us/frame, tile-shaped CTAs per SM {r.TileResidency()}, tile pairs {r.GetStats()['tile_pairs']}", flush=True)

t(mode)
if mode == "after":
    torch.zeros(1, device="cuda"); torch.cuda.synchronize()
    t("after one torch kernel")
