"""C3 frame time vs address layout: pre-allocate `pad` bytes before the context is created, print buffer addresses."""
import ctypes, sys
sys.path.insert(0, ".")
pad = int(sys.argv[1])
rt = ctypes.CDLL("libcudart.so.12")
if pad:
    p = ctypes.c_void_p(); rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(pad))
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
for _ in range(5): r.RenderMesh(m)
r.Synchronize()
r.TimerBegin()
for _ in range(30): r.RenderMesh(m)
t = r.TimerEnd() / 30 * 1e3
free, total = ctypes.c_size_t(), ctypes.c_size_t()
rt.cudaMemGetInfo(ctypes.byref(free), ctypes.byref(total))
print(f"pad {pad:>12d}: {t:8.1f} us/frame  color {r.DeviceColorPtr():#x} depth {r.DeviceDepthPtr():#x} used {(total.value-free.value)>>20} MiB")
