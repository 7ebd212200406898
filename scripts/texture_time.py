"""Textured shader: colour agreement with the oracle per filter, and frame time per filter at 1080p. GPU box only."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from edxraster_b200 import renderer as R, scenes
import parity

for f in range(6):
    sc = scenes.textured_plane(width=640, height=360, tex_filter=f)
    rep = parity.compare(parity.render_oracle(sc), parity.render_gpu(sc, stages=False))
    print(f"filter {f}: colour max diff {rep['color_max_diff']}, exact {100 * rep['color_exact_frac']:.3f} %", flush=True)

def timed(r, m, n=200):
    for _ in range(20): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(n): r.RenderMesh(m)
    return r.TimerEnd() / n

r = R.Renderer(0)
for name, mk in (("plane 1080p 512x512 tex", lambda: scenes.textured_plane(1920, 1080, quads=64, tex=(512, 512))),
                 ("C1 sphere 1080p, 3 slots", lambda: scenes.textured_sphere(1920, 1080, 100, 100))):
    sc = mk()
    r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
    m = r.CreateMesh(sc.vertices, sc.indices)
    r.SetPixelShader(2); base = timed(r, m)
    r.SetPixelShader(3); const = timed(r, m)
    m.SetTextures(sc["textures"], sc.get("tex_ids"))
    row = []
    for f in range(6):
        r.SetTextureFilter(f); row.append(timed(r, m))
    print(f"{name}: Lambert {base*1e3:.1f} us, constant albedo {const*1e3:.1f} us, textured by filter " + " ".join(f"{t*1e3:.1f}" for t in row) + " us", flush=True)
    m.Release()
for name in ("C1", "C3", "C4"):
    sc = scenes.by_name(name)
    r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
    m = r.CreateMesh(sc.vertices, sc.indices)
    print(f"{name} untextured: {timed(r, m, 100)*1e3:.1f} us", flush=True)
    m.Release()
