"""ctypes wrapper around oracle/_ref/libref.so: the reference's OWN sources (/root/reference/EDXRaster/Core, Utils),
unmodified, compiled by g++ against the EDXUtil stand-in in oracle/_ref_shim (recipe: oracle/Makefile, target `ref`).

TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's reference leg; nothing under edxraster_b200/ imports it.
/root/reference exists only in the build container, so the library is built there (`__graft_entry__.build()`) and
travels to the GPU box as a prebuilt, git-ignored file; `available()` says whether it is present.

One Renderer at a time: the reference keeps its state in a process-wide singleton (Core/RenderStates.h:35-44).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref.so")
_LIBS = {}
REFERENCE_SRC = "/root/reference/EDXRaster"

SHADER_BLINN_PHONG, SHADER_LAMBERT_ALBEDO = 1, 3


def build(force=False):
    """Compile the reference where its sources lie (no-op where /root/reference is absent)."""
    if not os.path.isdir(REFERENCE_SRC):
        return os.path.exists(_PATH)
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []))
    return True


def available():
    return os.path.exists(_PATH) or (os.path.isdir(REFERENCE_SRC) and build())


def _load(timing=False):
    path = _PATH.replace("libref.so", "libref_timing.so") if timing else _PATH
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, f32p, u32p, i32p, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.ref_create.restype = vp
    lib.ref_create.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.ref_destroy.argtypes = [vp]
    lib.ref_threads.argtypes = [vp]
    lib.ref_resize.argtypes = [vp, C.c_int, C.c_int]
    lib.ref_set_transform.argtypes = [vp, f32p, f32p, f32p]
    lib.ref_get_derived.argtypes = [vp, f32p, f32p]
    lib.ref_set_shader.argtypes = [vp, C.c_int]
    lib.ref_set_msaa.argtypes = [vp, C.c_int]
    lib.ref_set_hierarchical.argtypes = [vp, C.c_int]
    lib.ref_set_texture_filter.argtypes = [vp, C.c_int]
    lib.ref_samples.argtypes = [vp]
    lib.ref_set_mesh.argtypes = [vp, f32p, C.c_uint32, u32p, C.c_uint32, C.c_uint32, i32p, f32p, C.POINTER(u8p), i32p, u32p]
    lib.ref_render.restype = C.c_double
    lib.ref_render.argtypes = [vp]
    lib.ref_color.restype = u8p
    lib.ref_color.argtypes = [vp]
    lib.ref_get_color_sample.argtypes = [vp, C.c_int, u8p]
    lib.ref_get_depth_sample.argtypes = [vp, C.c_int, f32p]
    lib.ref_get_clip_verts.argtypes = [vp, f32p]
    lib.ref_num_raster_tris.restype = C.c_uint64
    lib.ref_num_raster_tris.argtypes = [vp]
    lib.ref_get_raster_tris.argtypes = [vp, i32p, f32p]
    lib.ref_get_winner_sample.argtypes = [vp, C.c_int, u32p]
    lib.ref_num_fragments.restype = C.c_uint64
    lib.ref_num_fragments.argtypes = [vp]
    _LIBS[path] = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Reference:
    """The reference Renderer (Core/Renderer.h:36-50) + Mesh (Utils/Mesh.h:29-68), driven in-process."""

    def __init__(self, width, height, threads=0, timing=False):
        self.lib = _load(timing)
        self.w, self.h = int(width), int(height)
        self.h_ = self.lib.ref_create(self.w, self.h, int(threads))
        if not self.h_:
            raise RuntimeError("a reference Renderer is already alive in this process (RenderStates is a singleton)")
        self.nv = 0

    def close(self):
        if self.h_:
            self.lib.ref_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return self.lib.ref_threads(self.h_)

    @property
    def samples(self):
        return self.lib.ref_samples(self.h_)

    def set_transform(self, model_view, proj, to_raster):
        mv, p, r = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (model_view, proj, to_raster))
        self.lib.ref_set_transform(self.h_, _p(mv, C.c_float), _p(p, C.c_float), _p(r, C.c_float))

    def derived(self):
        mvp = np.zeros(16, np.float32)
        eye = np.zeros(3, np.float32)
        self.lib.ref_get_derived(self.h_, _p(mvp, C.c_float), _p(eye, C.c_float))
        return mvp.reshape(4, 4), eye

    def set_shader(self, mode):
        """1 = BlinnPhongPixelShader (Shader.h:246-282), 3 = LambertianAlbedoPixelShader (:209-244, the default)."""
        if self.lib.ref_set_shader(self.h_, int(mode)) != 0:
            raise ValueError("the reference has no pixel shader for mode %d" % mode)

    def set_msaa(self, log2):
        self.lib.ref_set_msaa(self.h_, int(log2))

    def set_hierarchical(self, on):
        self.lib.ref_set_hierarchical(self.h_, 1 if on else 0)

    def set_texture_filter(self, f):
        self.lib.ref_set_texture_filter(self.h_, int(f))

    def set_mesh(self, vertices, indices, textures=None, tex_ids=None):
        """Mesh::LoadMesh (Utils/Mesh.cpp:11-34) fed from memory. `textures` as in orc.Oracle.set_textures; the default
        is the single constant 0.9-white slot LoadSphere / LoadPlane install (Mesh.cpp:47,66)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        textures = textures or [("constant", (0.9, 0.9, 0.9))]
        n = len(textures)
        kinds = np.zeros(n, np.int32)
        colors = np.zeros((n, 3), np.float32)
        dims = np.zeros((n, 2), np.int32)
        imgs = (C.POINTER(C.c_uint8) * n)()
        keep = []
        for k, (kind, val) in enumerate(textures):
            if kind == "constant":
                colors[k] = val
            else:
                img = np.ascontiguousarray(val, dtype=np.uint8)
                keep.append(img)
                kinds[k] = 1
                dims[k] = (img.shape[1], img.shape[0])
                imgs[k] = _p(img, C.c_uint8)
        ids = None if tex_ids is None else np.ascontiguousarray(tex_ids, dtype=np.uint32)
        self.nv = v.shape[0]
        self.lib.ref_set_mesh(self.h_, _p(v, C.c_float), v.shape[0], _p(i, C.c_uint32), i.shape[0], n, _p(kinds, C.c_int32),
                              _p(colors, C.c_float), imgs, _p(dims, C.c_int32), None if ids is None else _p(ids, C.c_uint32))

    def render(self):
        """Renderer::RenderMesh (Core/Renderer.cpp:100-118); returns the wall time in ms."""
        return float(self.lib.ref_render(self.h_))

    def color(self):
        """Renderer::GetBackBuffer (Renderer.cpp:360-363): H x W x 4 RGBA8, row 0 = bottom scanline."""
        buf = self.lib.ref_color(self.h_)
        return np.ctypeslib.as_array(buf, shape=(self.h, self.w, 4)).copy()

    def color_sample(self, sample=0):
        out = np.zeros((self.h, self.w, 4), np.uint8)
        self.lib.ref_get_color_sample(self.h_, int(sample), _p(out, C.c_uint8))
        return out

    def depth(self, sample=0):
        out = np.zeros((self.h, self.w), np.float32)
        self.lib.ref_get_depth_sample(self.h_, int(sample), _p(out, C.c_float))
        return out

    def winner_ordinal(self, sample=0):
        """Per pixel: position (in raster_tris order) of the triangle that wrote it last; 0xFFFFFFFF = none."""
        out = np.zeros((self.h, self.w), np.uint32)
        self.lib.ref_get_winner_sample(self.h_, int(sample), _p(out, C.c_uint32))
        return out

    def clip_verts(self):
        out = np.zeros((self.nv, 4), np.float32)
        self.lib.ref_get_clip_verts(self.h_, _p(out, C.c_float))
        return out

    def raster_tris(self):
        n = int(self.lib.ref_num_raster_tris(self.h_))
        ints = np.zeros((max(n, 1), 6), np.int32)
        flts = np.zeros((max(n, 1), 7), np.float32)
        self.lib.ref_get_raster_tris(self.h_, _p(ints, C.c_int32), _p(flts, C.c_float))
        return ints[:n], flts[:n]

    def num_fragments(self):
        return int(self.lib.ref_num_fragments(self.h_))
