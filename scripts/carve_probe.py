"""One-frame-at-a-time frame times of C3, C2, C4, C1 under edx_set_option("clip_carveout", v): `python scripts/carve_probe.py [torch] [v]`."""
import sys
sys.path.insert(0, ".")
carve = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 0
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch; torch.zeros(1, device="cuda")
from edxraster_b200 import renderer as R, scenes
out = []
for name in ("C3", "C2", "C4", "C1"):
    sc = scenes.by_name(name)
    r = R.Renderer(0)
    r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
    r.SetOption("clip_carveout", carve)
    m = r.CreateMesh(sc.vertices, sc.indices)
    for _ in range(10): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    n = 30 if name in ("C3", "C4") else 200
    for _ in range(n): r.RenderMesh(m)
    out.append(f"{name} {r.TimerEnd() / n * 1e3:.1f}")
    m.Release(); r.close()
print(" | ".join(out))
