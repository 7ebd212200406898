import sys, os, ctypes
sys.path.insert(0, ".")
mode = sys.argv[1]
rt = ctypes.CDLL("libcudart.so.12")
if mode == "lmemflag":
    print("cudaSetDeviceFlags ->", rt.cudaSetDeviceFlags(0x10 | 0x08))
from edxraster_b200 import renderer as R, scenes
sc = scenes.by_name("C3")
r = R.Renderer(0)
r.Initialize(sc.width, sc.height); r.SetTransform(sc.mv, sc.proj, sc.raster); r.SetPixelShader(sc.shader)
m = r.CreateMesh(sc.vertices, sc.indices)
fl = ctypes.c_uint(); rt.cudaGetDeviceFlags(ctypes.byref(fl)); print("flags %#x" % fl.value)
def t(tag):
    for _ in range(5): r.RenderMesh(m)
    r.Synchronize(); r.TimerBegin()
    for _ in range(30): r.RenderMesh(m)
    back = r.TimerEnd() / 30 * 1e3
    r.SetProfiling(True)
    acc = {"geom": 0, "clip": 0, "tile": 0, "total": 0}
    for _ in range(10):
        r.RenderMesh(m); r.Synchronize()
        st = r.GetStats()["stage_ms"]
        for k in acc: acc[k] += st[k] * 100
    r.SetProfiling(False)
    print(tag, f"back-to-back {back:.1f} us/frame | per-stage (synchronised) " + " ".join(f"{k} {v:.1f}" for k, v in acc.items()), flush=True)
t("before")
if mode == "nvrtc":
    from cuda import cuda, nvrtc
    src = b'extern "C" __global__ void k(int* p) { if (p) p[0] = 1; }'
    err, prog = nvrtc.nvrtcCreateProgram(src, b"k.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a"]
    print("compile", nvrtc.nvrtcCompileProgram(prog, len(opts), opts))
    err, sz = nvrtc.nvrtcGetCUBINSize(prog); cubin = b" " * sz; nvrtc.nvrtcGetCUBIN(prog, cubin)
    print("init", cuda.cuInit(0))
    err, mod = cuda.cuModuleLoadData(cubin); print("load", err)
    err, fn = cuda.cuModuleGetFunction(mod, b"k"); print("fn", err)
    import numpy as np
    arg = np.array([0], dtype=np.uint64); args = np.array([arg.ctypes.data], dtype=np.uint64)
    print("launch", cuda.cuLaunchKernel(fn, 1, 1, 1, 32, 1, 1, 0, 0, args.ctypes.data, 0))
    print("sync", cuda.cuCtxSynchronize())
    t("after foreign kernel")
