"""Minimal frame loop for ncu: renders `--frames` frames of one BASELINE config, nothing else."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edxraster_b200 import renderer as R, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--opt", action="append", default=[], help="name=value tuning option")
a = ap.parse_args()
sc = scenes.by_name(a.workload, a.scale)
r = R.Renderer(0)
r.Initialize(sc.width, sc.height)
r.SetTransform(sc.mv, sc.proj, sc.raster)
r.SetPixelShader(sc.shader)
for o in a.opt:
    k, v = o.split("=")
    r.SetOption(k, int(v))
m = r.CreateMesh(sc.vertices, sc.indices)
for _ in range(a.frames):
    r.RenderMesh(m)
    r.Synchronize()
print(a.workload, r.GetStats())
