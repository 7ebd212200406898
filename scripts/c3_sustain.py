"""Is the C3 frame time stable over long runs? (burst vs sustained clocks)"""
import sys, time, threading
sys.path.insert(0, ".")
import os
if os.environ.get("WITH_TORCH"):
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda")
import pynvml
from edxraster_b200 import renderer as R, scenes
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def clk(): return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
sc = scenes.by_name(sys.argv[1] if len(sys.argv) > 1 else "C3")
fresh = len(sys.argv) > 2
r = R.Renderer(0)
if not fresh:
    r.Initialize(1920, 1080)            # like a script that rendered something else first
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
ms_ = [r.CreateMesh(sc.vertices, sc.indices) for _ in range(int(os.environ.get("COPIES", "1")))]
for n in (3, 30, 30, 100, 300):
    r.Synchronize(); r.TimerBegin()
    for i in range(n): r.RenderMesh(ms_[i % len(ms_)])
    ms = r.TimerEnd() / n
    print(f"{n:5d} frames: {ms*1e3:8.1f} us/frame   sm clock {clk()[0]} MHz, {clk()[1]:.0f} W", flush=True)
print(r.GetStats())
