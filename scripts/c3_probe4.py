import sys, os, ctypes
sys.path.insert(0, ".")
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def state():
    return "sm %d mem %d MHz pstate %d power %.0f W" % (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
        pynvml.nvmlDeviceGetPerformanceState(h), pynvml.nvmlDeviceGetPowerUsage(h) / 1000)
from edxraster_b200 import renderer as R, scenes
rt = ctypes.CDLL("libcudart.so.12")
def limits():
    out = []
    for name, k in (("stack", 0), ("printf", 1), ("malloc_heap", 2), ("l2fetch", 5), ("persistL2", 6)):
        v = ctypes.c_size_t(); rt.cudaDeviceGetLimit(ctypes.byref(v), k); out.append("%s=%d" % (name, v.value))
    fl = ctypes.c_uint(); rt.cudaGetDeviceFlags(ctypes.byref(fl)); out.append("flags=%#x" % fl.value)
    return " ".join(out)
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
def t(tag):
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(60): r.RenderMesh(m)
    s = state()
    print(tag, f"{r.TimerEnd() / 60 * 1e3:.1f} us/frame |", s, "|", limits(), flush=True)
t("before")
mode = sys.argv[1]
if mode == "torch_tensor":
    import torch; x = torch.empty(1, device="cuda"); t("after torch.empty (malloc only)"); x.zero_(); torch.cuda.synchronize(); t("after torch kernel")
elif mode == "memset":
    p = ctypes.c_void_p(); rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(1 << 20)); rt.cudaMemset(p, 0, ctypes.c_size_t(1 << 20)); rt.cudaDeviceSynchronize(); t("after cudaMemset")
elif mode == "resize":
    r.Initialize(1920, 1080); r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); t("after re-Initialize")
