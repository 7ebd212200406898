"""ctypes wrapper around the CPU oracle (oracle/edx_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Nothing under edxraster_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile liborc.so / liborc_timing.so with the Makefile next to this file."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liborc.so", "liborc_timing.so"))
    src = os.path.join(_HERE, "edx_oracle.cpp")
    for n in ("liborc.so", "liborc_timing.so"):
        p = os.path.join(_HERE, n)
        if os.path.exists(p) and os.path.getmtime(p) < os.path.getmtime(src):
            need = True
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])


def _load(timing):
    name = "liborc_timing.so" if timing else "liborc.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, f32p, u32p, i32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_threads.argtypes = [vp]
    lib.orc_resize.argtypes = [vp, C.c_int, C.c_int]
    lib.orc_set_transform.argtypes = [vp, f32p, f32p, f32p]
    lib.orc_get_derived.argtypes = [vp, f32p, f32p, f32p]
    lib.orc_set_shader.argtypes = [vp, C.c_int]
    lib.orc_set_albedo.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    lib.orc_set_hierarchical.argtypes = [vp, C.c_int]
    u8p = C.POINTER(C.c_uint8)
    lib.orc_set_texture_filter.argtypes = [vp, C.c_int]
    lib.orc_clear_textures.argtypes = [vp]
    lib.orc_add_constant_texture.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    lib.orc_add_image_texture.argtypes = [vp, u8p, C.c_int, C.c_int]
    lib.orc_set_texture_ids.argtypes = [vp, u32p, C.c_uint32]
    lib.orc_tex_sample.argtypes = [vp, C.c_uint32, C.c_int] + [C.c_float] * 6 + [f32p]
    lib.orc_tex_levels.argtypes = [vp, C.c_uint32]
    lib.orc_tex_level.argtypes = [vp, C.c_uint32, C.c_int, u8p]
    lib.orc_set_msaa.argtypes = [vp, C.c_int]
    lib.orc_samples.argtypes = [vp]
    lib.orc_get_winner_sample.argtypes = [vp, C.c_int, u32p]
    lib.orc_get_depth_sample.argtypes = [vp, C.c_int, f32p]
    lib.orc_render.argtypes = [vp, f32p, C.c_uint32, u32p, C.c_uint32]
    lib.orc_color.restype = C.POINTER(C.c_uint8)
    lib.orc_color.argtypes = [vp]
    lib.orc_get_winner.argtypes = [vp, u32p]
    lib.orc_get_depth.argtypes = [vp, f32p]
    lib.orc_get_clip_verts.argtypes = [vp, f32p]
    lib.orc_num_raster_tris.restype = C.c_uint64
    lib.orc_num_raster_tris.argtypes = [vp]
    lib.orc_get_raster_tris.argtypes = [vp, i32p, f32p]
    lib.orc_get_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.orc_snap.argtypes = [C.c_float]
    lib.orc_snap.restype = C.c_int
    lib.orc_clip_code.argtypes = [C.c_float] * 4
    lib.orc_clip_code.restype = C.c_uint32
    lib.orc_clip_triangle.argtypes = [f32p, f32p, f32p]
    lib.orc_mat_mul.argtypes = [f32p, f32p, f32p]
    lib.orc_mat_inverse.argtypes = [f32p, f32p]
    _LIBS[name] = lib
    return lib


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


SHADER_DEPTH_ONLY, SHADER_BLINN_PHONG, SHADER_LAMBERT, SHADER_LAMBERT_ALBEDO = 0, 1, 2, 3


class Oracle:
    """Mirror of the reference Renderer (Core/Renderer.h:36-50) on the CPU restatement."""

    def __init__(self, width, height, threads=0, timing=False):
        self.lib = _load(timing)
        self.w, self.h = int(width), int(height)
        self.h_ = self.lib.orc_create(self.w, self.h, int(threads))
        self.nv = 0

    def close(self):
        if self.h_:
            self.lib.orc_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return self.lib.orc_threads(self.h_)

    def set_transform(self, model_view, proj, to_raster):
        mv, p, r = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (model_view, proj, to_raster))
        self.lib.orc_set_transform(self.h_, _f32(mv), _f32(p), _f32(r))

    def derived(self):
        mvp = np.zeros(16, np.float32)
        eye = np.zeros(3, np.float32)
        light = np.zeros(3, np.float32)
        self.lib.orc_get_derived(self.h_, _f32(mvp), _f32(eye), _f32(light))
        return mvp.reshape(4, 4), eye, light

    def set_shader(self, mode):
        self.lib.orc_set_shader(self.h_, int(mode))

    def set_albedo(self, r, g, b):
        self.lib.orc_set_albedo(self.h_, r, g, b)

    def set_texture_filter(self, f):
        """Renderer::SetTextureFilter (Core/Renderer.h:48): 0 nearest, 1 linear, 2 trilinear, 3/4/5 anisotropic 4x/8x/16x."""
        self.lib.orc_set_texture_filter(self.h_, int(f))

    def set_textures(self, textures, tex_ids=None):
        """Mesh::mTextures + GetTextureIds (Utils/Mesh.h:23,54-59). `textures`: list of ('constant', (r, g, b)) or
        ('image', uint8 array H x W x 4); `tex_ids`: one slot per submitted triangle (None = all 0)."""
        self.lib.orc_clear_textures(self.h_)
        self._tex_dims = []
        for kind, val in textures:
            if kind == "constant":
                self.lib.orc_add_constant_texture(self.h_, float(val[0]), float(val[1]), float(val[2]))
                self._tex_dims.append(None)
            else:
                img = np.ascontiguousarray(val, dtype=np.uint8)
                assert img.ndim == 3 and img.shape[2] == 4
                self.lib.orc_add_image_texture(self.h_, img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[1], img.shape[0])
                self._tex_dims.append((img.shape[1], img.shape[0]))
        if tex_ids is not None:
            ids = np.ascontiguousarray(tex_ids, dtype=np.uint32)
            self.lib.orc_set_texture_ids(self.h_, ids.ctypes.data_as(C.POINTER(C.c_uint32)), ids.shape[0])

    def tex_sample(self, slot, filt, u, v, d0=(0.0, 0.0), d1=(0.0, 0.0)):
        out = np.zeros(3, np.float32)
        self.lib.orc_tex_sample(self.h_, slot, filt, u, v, d0[0], d0[1], d1[0], d1[1], _f32(out))
        return out

    def tex_mips(self, slot):
        w, h = self._tex_dims[slot]
        out = []
        for l in range(self.lib.orc_tex_levels(self.h_, slot)):
            lw, lh = max(1, w >> l), max(1, h >> l)
            a = np.zeros((lh, lw, 4), np.uint8)
            self.lib.orc_tex_level(self.h_, slot, l, a.ctypes.data_as(C.POINTER(C.c_uint8)))
            out.append(a)
        return out

    def set_msaa(self, log2):
        """Renderer::SetMSAAMode (Core/Renderer.cpp:94-98): 2^log2 samples per pixel."""
        self.lib.orc_set_msaa(self.h_, int(log2))

    @property
    def samples(self):
        return self.lib.orc_samples(self.h_)

    def set_hierarchical(self, on):
        self.lib.orc_set_hierarchical(self.h_, 1 if on else 0)

    def render(self, vertices, indices):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        self._keep = (v, i)
        self.nv = v.shape[0]
        self.lib.orc_render(self.h_, _f32(v), v.shape[0], _u32(i), i.shape[0])

    def color(self):
        buf = self.lib.orc_color(self.h_)
        return np.ctypeslib.as_array(buf, shape=(self.h, self.w, 4)).copy()

    def depth(self, sample=0):
        out = np.zeros((self.h, self.w), np.float32)
        self.lib.orc_get_depth_sample(self.h_, int(sample), _f32(out))
        return out

    def winner(self, sample=0):
        out = np.zeros((self.h, self.w), np.uint32)
        self.lib.orc_get_winner_sample(self.h_, int(sample), _u32(out))
        return out

    def clip_verts(self):
        out = np.zeros((self.nv, 4), np.float32)
        self.lib.orc_get_clip_verts(self.h_, _f32(out))
        return out

    def raster_tris(self):
        n = int(self.lib.orc_num_raster_tris(self.h_))
        ints = np.zeros((max(n, 1), 7), np.int32)
        flts = np.zeros((max(n, 1), 7), np.float32)
        self.lib.orc_get_raster_tris(self.h_, ints.ctypes.data_as(C.POINTER(C.c_int32)), _f32(flts))
        return ints[:n], flts[:n]

    def stats(self):
        c = (C.c_uint64 * 4)()
        ms = (C.c_double * 6)()
        self.lib.orc_get_stats(self.h_, c, ms)
        names = ("vertex", "clip_setup", "bin", "raster", "shade", "fb_update")
        return {"raster_tris": int(c[0]), "fragments": int(c[1]), "covered_samples": int(c[2]), "bin_refs": int(c[3]),
                "ms": {k: float(v) for k, v in zip(names, ms)}}


def snap(f):
    return int(_load(False).orc_snap(float(f)))


def clip_code(x, y, z, w):
    return int(_load(False).orc_clip_code(x, y, z, w))


def clip_triangle(tri):
    t = np.ascontiguousarray(tri, dtype=np.float32).reshape(12)
    pos = np.zeros((16, 4), np.float32)
    wt = np.zeros((16, 3), np.float32)
    n = _load(False).orc_clip_triangle(_f32(t), _f32(pos), _f32(wt))
    if n < 0:
        return None, None
    return pos[:n].copy(), wt[:n].copy()


def mat_mul(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(16)
    b = np.ascontiguousarray(b, np.float32).reshape(16)
    o = np.zeros(16, np.float32)
    _load(False).orc_mat_mul(_f32(a), _f32(b), _f32(o))
    return o.reshape(4, 4)


def mat_inverse(a):
    a = np.ascontiguousarray(a, np.float32).reshape(16)
    o = np.zeros(16, np.float32)
    _load(False).orc_mat_inverse(_f32(a), _f32(o))
    return o.reshape(4, 4)
