"""Per-config frame and stage times under a set of tuning options (one frame at a time and 3 in flight).
usage: python scripts/stage_bench.py [--workloads C1,C2,C3,C4] [--opt name=value ...] [--label text]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="C1,C2,C3,C4")
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--variants", default="", help="semicolon-separated option sets, e.g. 'mid_max=0,small_max=32;mid_max=64'")
a = ap.parse_args()
variants = [v for v in a.variants.split(";") if v] or [",".join(a.opt)]
for w in a.workloads.split(","):
    sc = scenes.by_name(w, 1.0)
    for var in variants:
        opts = dict(kv.split("=") for kv in var.split(",") if kv)
        r = R.Renderer(0)
        r.Initialize(sc.width, sc.height)
        r.SetTransform(sc.mv, sc.proj, sc.raster)
        r.SetPixelShader(sc.shader)
        for k, v in opts.items():
            r.SetOption(k, int(v))
        m = r.CreateMesh(sc.vertices, sc.indices)
        frames = 30 if sc.num_tris < 5_000_000 else 10
        r.RenderMesh(m)
        r.Synchronize()                  # the first frame sizes the internal queues (grow + re-run if it overflowed)
        for _ in range(3):
            r.RenderMesh(m)
        r.Synchronize()
        r.TimerBegin()
        for _ in range(frames):
            r.RenderMesh(m)
        ms = r.TimerEnd() / frames
        r.SetProfiling(True)
        st = {"geom": 0.0, "clip": 0.0, "tile": 0.0}
        for _ in range(5):
            r.RenderMesh(m)
            r.Synchronize()
            s = r.GetStats()
            for k in st:
                st[k] += s["stage_ms"][k] / 5
        r.SetProfiling(False)
        s = r.GetStats()
        ring = R.FrameRing(0, depth=3)
        ring.Initialize(sc.width, sc.height)
        ring.SetPixelShader(sc.shader)
        for k, v in opts.items():
            ring.SetOption(k, int(v))
        xf = R.PackedTransform(sc.mv, sc.proj, sc.raster)
        n = frames * 4
        for lane in ring.lanes:          # size each lane's queues before frames pile up unsynchronised
            lane.SetTransform(xf)
            lane.RenderMesh(m)
            lane.Synchronize()
        for rep in range(2):
            ring.Synchronize()
            t = time.perf_counter()
            for i in range(n):
                ring.Submit(m, xf)
            ring.Synchronize()
            ms3 = (time.perf_counter() - t) * 1e3 / n
        ring.close()
        print("%s [%s] frame %.1f us | 3 in flight %.1f us | geom %.1f clip %.1f tile %.1f | mid %d big %d clipped %d pairs %d listpairs %d"
              % (w, var, ms * 1e3, ms3 * 1e3, st["geom"] * 1e3, st["clip"] * 1e3, st["tile"] * 1e3, s["mid_tris"], s["binned_tris"], s["clipped_tris"], s["tile_pairs"], s["bin_pairs"]), flush=True)
        m.Release()
        r.close()
