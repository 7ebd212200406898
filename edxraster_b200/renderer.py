"""Python mirror of the reference's renderer-facing API over the C ABI.

Same member names and argument meaning as EDX::RasterRenderer::Renderer (Core/Renderer.h:36-50) and
Mesh (Utils/Mesh.h:29-68) so parity tests read like calls into the reference. Everything executes in
the CUDA library; this file only marshals numpy arrays through ctypes.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EdxError

SHADER_DEPTH_ONLY, SHADER_BLINN_PHONG, SHADER_LAMBERT, SHADER_LAMBERT_ALBEDO = 0, 1, 2, 3


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class PackedTransform:
    """The three SetTransform matrices already marshalled for the C ABI (row-major float[16] each). Passing
    one to SetTransform / FrameRing.Submit skips ~15 us of numpy conversion per call, which matters to a
    host loop that submits a frame every 45 us."""
    __slots__ = ("mv", "proj", "raster")

    def __init__(self, model_view, proj, to_raster):
        conv = lambda m: (C.c_float * 16)(*np.asarray(m, dtype=np.float32).reshape(16).tolist())
        self.mv, self.proj, self.raster = conv(model_view), conv(proj), conv(to_raster)


class Mesh:
    """Utils/Mesh.h: owns the vertex / index buffers (here: device copies)."""

    def __init__(self, renderer, vertices, indices):
        self._r = renderer
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        self.num_verts, self.num_tris = v.shape[0], i.shape[0]
        h = C.c_void_p()
        renderer._check(renderer._lib.edx_mesh_create(renderer._h, v.ctypes.data, v.shape[0], i.ctypes.data, i.shape[0], None, C.byref(h)))
        self._h = h

    def update(self, vertices_ptr, num_verts, indices_ptr, num_tris):
        """Re-upload from raw host pointers (e.g. pinned torch tensors); asynchronous."""
        self._r._check(self._r._lib.edx_mesh_update(self._r._h, self._h, vertices_ptr, num_verts, indices_ptr, num_tris))
        self.num_verts, self.num_tris = num_verts, num_tris

    def SetTextures(self, textures, tex_ids=None):
        """Mesh::mTextures + GetTextureIds (Utils/Mesh.h:23,54-59). `textures`: list of ('constant', (r, g, b)) or
        ('image', uint8 array H x W x 4, row 0 at v = 0); `tex_ids`: one slot per triangle (None = all 0)."""
        descs = (_lib.TextureDesc * max(1, len(textures)))()
        keep = []
        for d, (kind, val) in zip(descs, textures):
            if kind == "constant":
                d.kind = 0
                d.color[0], d.color[1], d.color[2] = float(val[0]), float(val[1]), float(val[2])
            else:
                img = np.ascontiguousarray(val, dtype=np.uint8)
                if img.ndim != 3 or img.shape[2] != 4:
                    raise ValueError("image textures are H x W x 4 uint8")
                keep.append(img)
                d.kind, d.rgba8, d.width, d.height = 1, img.ctypes.data, img.shape[1], img.shape[0]
        ids = None
        if tex_ids is not None:
            ids = np.ascontiguousarray(tex_ids, dtype=np.uint32)
            if ids.shape[0] != self.num_tris:
                raise ValueError("one texture id per triangle")
        self._r._check(self._r._lib.edx_mesh_set_textures(self._r._h, self._h, descs, len(textures), None if ids is None else ids.ctypes.data))

    def TextureLevel(self, slot, level):
        """one mip level of an image texture, read back from the device (diagnostics)"""
        w, h = C.c_uint32(), C.c_uint32()
        self._r._check(self._r._lib.edx_mesh_read_texture_level(self._r._h, self._h, slot, level, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value, 4), np.uint8)
        self._r._check(self._r._lib.edx_mesh_read_texture_level(self._r._h, self._h, slot, level, out.ctypes.data, None, None))
        return out

    def Release(self):
        if self._h:
            self._r._lib.edx_mesh_destroy(self._r._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.Release()
        except Exception:
            pass


class Renderer:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.edx_create(int(device), C.byref(h))
        if rc != 0:
            raise EdxError(rc, "edx_create failed (no B200 / CUDA device?)")
        self._h = h
        self.width = self.height = 0

    def _check(self, rc):
        if rc != 0:
            raise EdxError(rc, self._lib.edx_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.edx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- Core/Renderer.h:40-50 -------------------------------------------------------------
    def Initialize(self, width, height):
        self._check(self._lib.edx_initialize(self._h, width, height))
        self.width, self.height = int(width), int(height)

    def Resize(self, width, height):
        self._check(self._lib.edx_resize(self._h, width, height))
        self.width, self.height = int(width), int(height)

    def SetTransform(self, model_view, proj=None, to_raster=None):
        if isinstance(model_view, PackedTransform):
            t = model_view
            self._check(self._lib.edx_set_transform(self._h, t.mv, t.proj, t.raster))
            return
        mv, p, r = (np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in (model_view, proj, to_raster))
        self._check(self._lib.edx_set_transform(self._h, _f32(mv), _f32(p), _f32(r)))

    def RenderMesh(self, mesh):
        self._check(self._lib.edx_render_mesh(self._h, mesh._h))

    def GetBackBuffer(self):
        p = self._lib.edx_get_back_buffer(self._h)
        if not p:
            raise EdxError(-2, self._lib.edx_last_error(self._h).decode())
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 4))

    def SetMSAAMode(self, log2):
        self._check(self._lib.edx_set_msaa_mode(self._h, log2))

    def SetTextureFilter(self, f):
        self._check(self._lib.edx_set_texture_filter(self._h, f))

    def SetHierarchicalRasterize(self, on):
        self._check(self._lib.edx_set_hierarchical_rasterize(self._h, 1 if on else 0))

    def WriteFrameToFile(self, path):
        self._check(self._lib.edx_write_frame_to_file(self._h, path.encode()))

    # --- extensions -------------------------------------------------------------------------
    def SetPixelShader(self, shader):
        self._check(self._lib.edx_set_pixel_shader(self._h, shader))

    def SetAlbedo(self, r, g, b):
        self._check(self._lib.edx_set_albedo(self._h, r, g, b))

    def CreateMesh(self, vertices, indices):
        return Mesh(self, vertices, indices)

    def Synchronize(self):
        self._check(self._lib.edx_synchronize(self._h))

    def GetDepthBuffer(self):
        out = np.empty((self.height, self.width), np.float32)
        self._check(self._lib.edx_read_depth(self._h, _f32(out)))
        return out

    def SetCaptureIds(self, on):
        self._check(self._lib.edx_set_capture_ids(self._h, 1 if on else 0))

    def GetWinnerIds(self):
        out = np.empty((self.height, self.width), np.uint32)
        self._check(self._lib.edx_read_winner_ids(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def GetSample(self, sample):
        """(depth, owner ids) of one MSAA sample plane."""
        d = np.empty((self.height, self.width), np.float32)
        i = np.empty((self.height, self.width), np.uint32)
        self._check(self._lib.edx_read_sample(self._h, int(sample), _f32(d), i.ctypes.data_as(C.POINTER(C.c_uint32))))
        return d, i

    def DebugClipVertices(self, mesh):
        out = np.empty((mesh.num_verts, 4), np.float32)
        self._check(self._lib.edx_debug_clip_vertices(self._h, mesh._h, _f32(out)))
        return out

    def DebugRasterTriangles(self, mesh, capacity=None):
        cap = int(capacity or (mesh.num_tris * 7 + 16))
        ints = np.zeros((cap, 7), np.int32)
        flts = np.zeros((cap, 7), np.float32)
        n = C.c_uint64(0)
        self._check(self._lib.edx_debug_raster_triangles(self._h, mesh._h, cap, ints.ctypes.data_as(C.POINTER(C.c_int32)), _f32(flts), C.byref(n)))
        return ints[:n.value], flts[:n.value]

    def DerivedState(self):
        mvp, eye, light = np.zeros(16, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        self._check(self._lib.edx_get_derived_state(self._h, _f32(mvp), _f32(eye), _f32(light)))
        return mvp.reshape(4, 4), eye, light

    def DeviceColorPtr(self):
        return self._lib.edx_device_color(self._h)

    def DeviceDepthPtr(self):
        return self._lib.edx_device_depth(self._h)

    def SetRenderTarget(self, color_ptr, depth_ptr):
        self._check(self._lib.edx_set_render_target(self._h, C.c_void_p(color_ptr or None), C.c_void_p(depth_ptr or None)))

    def SetFrameSink(self, color_ptr, depth_ptr):
        """After every frame, push the finished colour / depth buffer to these device addresses (own or peer-mapped
        memory) with the copy engine, stream-ordered behind the frame: the frame-parallel gather."""
        self._check(self._lib.edx_set_frame_sink(self._h, C.c_void_p(color_ptr or None), C.c_void_p(depth_ptr or None)))

    def DeviceAlloc(self, nbytes):
        """Device memory on this context's GPU (e.g. a frame store for SetFrameSink); returns the address."""
        out = C.c_void_p()
        self._check(self._lib.edx_device_alloc(self._h, C.c_size_t(nbytes), C.byref(out)))
        return out.value

    def DeviceFree(self, ptr):
        self._check(self._lib.edx_device_free(self._h, C.c_void_p(ptr)))

    def ReadDevice(self, ptr, nbytes):
        """Stream-ordered copy of device memory to a new numpy byte array (synchronises)."""
        host = np.empty(nbytes, dtype=np.uint8)
        self._check(self._lib.edx_read_device(self._h, host.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(nbytes)))
        return host

    def SetFrameSinkSignal(self, word_ptr):
        """Behind every frame's pushes the context stores the number of frames pushed so far to this device address."""
        self._check(self._lib.edx_set_frame_sink_signal(self._h, C.c_void_p(word_ptr or None)))

    def FlushFrameSink(self):
        """The context's stream waits (on the device) for every push issued so far."""
        self._check(self._lib.edx_flush_frame_sink(self._h))

    def ReadDepthInto(self, host_ptr):
        """Copy the depth buffer to caller memory (pinned for full PCIe speed); synchronises."""
        self._check(self._lib.edx_read_depth(self._h, C.cast(C.c_void_p(host_ptr), C.POINTER(C.c_float))))

    def SetScreenPartition(self, part, parts):
        """Sort-first split: own only the 64x64 bins b with b % parts == part."""
        self._check(self._lib.edx_set_screen_partition(self._h, int(part), int(parts)))

    def SetStream(self, cuda_stream_ptr):
        self._check(self._lib.edx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def TimerBegin(self):
        self._check(self._lib.edx_timer_begin(self._h))

    def TimerEnd(self):
        ms = C.c_float(0)
        self._check(self._lib.edx_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    def SetProfiling(self, on):
        self._check(self._lib.edx_set_profiling(self._h, 1 if on else 0))

    def GetStats(self):
        s = _lib.Stats()
        self._check(self._lib.edx_get_stats(self._h, C.byref(s)))
        return {"submitted_tris": s.submitted_tris, "clipped_tris": s.clipped_tris, "binned_tris": s.binned_tris,
                "clip_records": s.clip_records, "regrow_count": s.regrow_count, "tile_pairs": s.tile_pairs, "mid_tris": s.mid_tris, "bin_pairs": s.bin_pairs,
                "stage_ms": {"geom": s.stage_ms[0], "clip": s.stage_ms[1], "tile": s.stage_ms[2], "total": s.stage_ms[3]}}

    def SetOption(self, name, value):
        self._check(self._lib.edx_set_option(self._h, name.encode(), int(value)))

    def TileResidency(self):
        n = C.c_int(0)
        self._check(self._lib.edx_debug_tile_residency(self._h, C.byref(n)))
        return n.value

    def LastLaunchCount(self):
        return int(self._lib.edx_last_launch_count(self._h))

    def LastLaunchList(self):
        """Names of the kernels the last RenderMesh launched, in order."""
        s = self._lib.edx_last_launch_list(self._h)
        return [k for k in (s.decode() if s else "").split(",") if k]


class FrameRing:
    """Several frames in flight on one GPU. A frame is three dependent kernels of very different shapes (a
    chip-wide geometry pass, a short clipper, one CTA per screen bin); alone they leave most of the B200 idle
    between and inside them, so rendering throughput rises 1.3-3x when 3-4 independent frames overlap
    (DESIGN.md section 7). A ring is `depth` contexts on their own streams sharing the meshes; Submit() rotates
    over them and returns a ticket, GetBackBuffer(ticket) / ReadDepthInto(ticket, ptr) waits for that frame only.
    A ticket stays valid until `depth` more frames have been submitted. The reference renders one frame at a
    time (Core/Renderer.cpp:100-118); this is the throughput form of the same call for frame farms."""

    def __init__(self, device=0, depth=3):
        if depth < 1:
            raise ValueError("depth >= 1")
        self.lanes = [Renderer(device) for _ in range(depth)]
        self.depth = depth
        self._ticket = 0

    def _all(self, name, *a):
        for r in self.lanes:
            getattr(r, name)(*a)

    def Initialize(self, width, height): self._all("Initialize", width, height)
    def Resize(self, width, height): self._all("Resize", width, height)
    def SetPixelShader(self, shader): self._all("SetPixelShader", shader)
    def SetAlbedo(self, r, g, b): self._all("SetAlbedo", r, g, b)
    def SetMSAAMode(self, log2): self._all("SetMSAAMode", log2)
    def SetTextureFilter(self, f): self._all("SetTextureFilter", f)
    def SetHierarchicalRasterize(self, on): self._all("SetHierarchicalRasterize", on)
    def SetOption(self, name, value): self._all("SetOption", name, value)
    def Synchronize(self): self._all("Synchronize")

    def CreateMesh(self, vertices, indices):
        """uploaded once (by lane 0, synchronously); every lane may render it"""
        return self.lanes[0].CreateMesh(vertices, indices)

    def Submit(self, mesh, model_view, proj=None, to_raster=None, color_ptr=None, depth_ptr=None):
        """SetTransform + RenderMesh on the next lane; optional caller-owned device targets for this frame.
        `model_view` may be a PackedTransform (then proj / to_raster are omitted)."""
        t = self._ticket
        r = self.lanes[t % self.depth]
        if color_ptr is not None or depth_ptr is not None:
            r.SetRenderTarget(color_ptr or 0, depth_ptr or 0)
        r.SetTransform(model_view, proj, to_raster)
        r.RenderMesh(mesh)
        self._ticket += 1
        return t

    def _lane(self, ticket):
        if not (self._ticket - self.depth <= ticket < self._ticket) or ticket < 0:
            raise ValueError("ticket %d is no longer in the ring" % ticket)
        return self.lanes[ticket % self.depth]

    def Wait(self, ticket): self._lane(ticket).Synchronize()
    def GetBackBuffer(self, ticket): return self._lane(ticket).GetBackBuffer()
    def GetDepthBuffer(self, ticket): return self._lane(ticket).GetDepthBuffer()
    def ReadDepthInto(self, ticket, host_ptr): self._lane(ticket).ReadDepthInto(host_ptr)

    def close(self):
        for r in self.lanes:
            r.close()
