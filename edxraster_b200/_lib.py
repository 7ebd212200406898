"""ctypes loader for the product library (libedxraster_b200.so, built from csrc/).

There is no fallback: if the library is missing or no B200 is present, importing callers get an
exception (LibraryMissing / EdxError), never a CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EDX_LIB: load another build of the same library (the diagnostic `make -C csrc debug` build); never a different implementation
LIB_PATH = os.environ.get("EDX_LIB") or os.path.join(_HERE, "libedxraster_b200.so")

# every symbol include/edxraster_c.h declares (tests check the library exports all of them)
SYMBOLS = [
    "edx_create", "edx_destroy", "edx_last_error", "edx_version", "edx_initialize", "edx_resize",
    "edx_set_transform", "edx_set_msaa_mode", "edx_set_texture_filter", "edx_set_hierarchical_rasterize",
    "edx_write_frame_to_file", "edx_set_pixel_shader", "edx_set_albedo", "edx_mesh_create", "edx_mesh_update",
    "edx_mesh_destroy", "edx_render_mesh", "edx_get_back_buffer", "edx_synchronize", "edx_read_depth",
    "edx_set_capture_ids", "edx_read_winner_ids", "edx_read_sample", "edx_debug_clip_vertices", "edx_debug_raster_triangles",
    "edx_get_derived_state", "edx_device_color", "edx_device_depth", "edx_set_render_target", "edx_set_frame_sink", "edx_set_frame_sink_signal", "edx_flush_frame_sink", "edx_device_count", "edx_enable_peer_access", "edx_device_alloc", "edx_device_free", "edx_read_device", "edx_set_screen_partition", "edx_set_stream", "edx_timer_begin",
    "edx_timer_end", "edx_set_profiling", "edx_get_stats", "edx_set_option", "edx_last_launch_count", "edx_last_launch_list",
    "edx_mesh_set_textures", "edx_mesh_read_texture_level", "edx_debug_tile_residency",
]

EDX_OK, EDX_ERR_INVALID, EDX_ERR_CUDA, EDX_ERR_OOM, EDX_ERR_OVERFLOW, EDX_ERR_UNSUPPORTED, EDX_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6


class LibraryMissing(RuntimeError):
    pass


class EdxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("edxraster error %d: %s" % (code, msg))
        self.code = code


class TextureDesc(C.Structure):         # edx_texture_desc
    _fields_ = [("kind", C.c_int), ("color", C.c_float * 3), ("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("submitted_tris", C.c_uint64), ("clipped_tris", C.c_uint64), ("binned_tris", C.c_uint64),
                ("clip_records", C.c_uint64), ("regrow_count", C.c_uint32), ("tile_pairs", C.c_uint32),
                ("mid_tris", C.c_uint64), ("bin_pairs", C.c_uint64), ("stage_ms", C.c_float * 8)]


_lib = None


def load():
    """Load the C-ABI library and declare its prototypes. Raises LibraryMissing if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "or `make -C edxraster_b200/csrc` (no CPU fallback exists)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, f32p, u32p, i32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    lib.edx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.edx_destroy.argtypes = [vp]
    lib.edx_destroy.restype = None
    lib.edx_last_error.argtypes = [vp]
    lib.edx_last_error.restype = C.c_char_p
    lib.edx_version.restype = C.c_char_p
    lib.edx_initialize.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.edx_resize.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.edx_set_transform.argtypes = [vp, f32p, f32p, f32p]
    lib.edx_set_msaa_mode.argtypes = [vp, C.c_int]
    lib.edx_set_texture_filter.argtypes = [vp, C.c_int]
    lib.edx_set_hierarchical_rasterize.argtypes = [vp, C.c_int]
    lib.edx_write_frame_to_file.argtypes = [vp, C.c_char_p]
    lib.edx_set_pixel_shader.argtypes = [vp, C.c_int]
    lib.edx_set_albedo.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    lib.edx_mesh_create.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.POINTER(vp)]
    lib.edx_mesh_update.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_uint32]
    lib.edx_mesh_destroy.argtypes = [vp, vp]
    lib.edx_render_mesh.argtypes = [vp, vp]
    lib.edx_get_back_buffer.argtypes = [vp]
    lib.edx_get_back_buffer.restype = C.POINTER(C.c_uint8)
    lib.edx_synchronize.argtypes = [vp]
    lib.edx_read_depth.argtypes = [vp, f32p]
    lib.edx_set_capture_ids.argtypes = [vp, C.c_int]
    lib.edx_read_winner_ids.argtypes = [vp, u32p]
    lib.edx_read_sample.argtypes = [vp, C.c_int, f32p, u32p]
    lib.edx_debug_clip_vertices.argtypes = [vp, vp, f32p]
    lib.edx_debug_raster_triangles.argtypes = [vp, vp, C.c_uint64, i32p, f32p, C.POINTER(C.c_uint64)]
    lib.edx_get_derived_state.argtypes = [vp, f32p, f32p, f32p]
    lib.edx_device_color.argtypes = [vp]
    lib.edx_device_color.restype = vp
    lib.edx_device_depth.argtypes = [vp]
    lib.edx_device_depth.restype = vp
    lib.edx_set_render_target.argtypes = [vp, vp, vp]
    lib.edx_set_frame_sink.argtypes = [vp, vp, vp]
    lib.edx_set_frame_sink_signal.argtypes = [vp, vp]
    lib.edx_flush_frame_sink.argtypes = [vp]
    lib.edx_enable_peer_access.argtypes = [vp, C.c_int]
    lib.edx_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.edx_device_free.argtypes = [vp, vp]
    lib.edx_read_device.argtypes = [vp, vp, vp, C.c_size_t]
    lib.edx_set_screen_partition.argtypes = [vp, C.c_int, C.c_int]
    lib.edx_set_stream.argtypes = [vp, vp]
    lib.edx_timer_begin.argtypes = [vp]
    lib.edx_timer_end.argtypes = [vp, f32p]
    lib.edx_set_profiling.argtypes = [vp, C.c_int]
    lib.edx_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.edx_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.edx_last_launch_count.argtypes = [vp]
    lib.edx_last_launch_list.argtypes = [vp]
    lib.edx_last_launch_list.restype = C.c_char_p
    lib.edx_mesh_set_textures.argtypes = [vp, vp, C.POINTER(TextureDesc), C.c_uint32, vp]
    lib.edx_debug_tile_residency.argtypes = [vp, C.POINTER(C.c_int)]
    lib.edx_mesh_read_texture_level.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, u32p, u32p]
    _lib = lib
    return lib
