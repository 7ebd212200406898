"""The C-ABI library loads and exports exactly what include/edxraster_c.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from edxraster_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "edxraster_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edx_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SYMBOLS) == names


def test_version_names_the_architecture():
    assert b"sm_100a" in _lib.load().edx_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert _lib.load().edx_create(0, C.byref(h)) == _lib.EDX_ERR_NO_DEVICE
    assert not h.value
    from edxraster_b200 import renderer
    with pytest.raises(_lib.EdxError):
        renderer.Renderer(0)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (only tests, smoke and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "edxraster_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle,", "") or fn == "scenes.py", fn
    for fn in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", fn)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()


def test_null_handles_are_rejected_not_dereferenced():
    """Every entry point given NULL handles returns an error status (or NULL) instead of crashing - 'never abort'
    (errors). Runs without a GPU: nothing here reaches CUDA."""
    lib = _lib.load()
    f16 = (C.c_float * 16)()
    buf = (C.c_float * 64)()
    u = C.c_uint32(0)
    n64 = C.c_uint64(0)
    ms = C.c_float(0)
    st = _lib.Stats()
    tex = (_lib.TextureDesc * 1)()
    out = C.c_void_p()
    calls = {
        "edx_initialize": (None, 64, 64), "edx_resize": (None, 64, 64), "edx_set_transform": (None, f16, f16, f16),
        "edx_set_msaa_mode": (None, 1), "edx_set_texture_filter": (None, 1), "edx_set_hierarchical_rasterize": (None, 1),
        "edx_write_frame_to_file": (None, b"/tmp/x.bmp"), "edx_set_pixel_shader": (None, 1), "edx_set_albedo": (None, 0.5, 0.5, 0.5),
        "edx_mesh_create": (None, None, 0, None, 0, None, C.byref(out)), "edx_mesh_update": (None, None, None, 0, None, 0),
        "edx_mesh_set_textures": (None, None, tex, 1, None), "edx_mesh_read_texture_level": (None, None, 0, 0, None, C.byref(u), C.byref(u)),
        "edx_render_mesh": (None, None), "edx_synchronize": (None,), "edx_read_depth": (None, buf),
        "edx_set_capture_ids": (None, 1), "edx_read_winner_ids": (None, C.cast(buf, C.POINTER(C.c_uint32))),
        "edx_read_sample": (None, 0, buf, C.cast(buf, C.POINTER(C.c_uint32))), "edx_debug_clip_vertices": (None, None, buf),
        "edx_debug_raster_triangles": (None, None, 0, C.cast(buf, C.POINTER(C.c_int32)), buf, C.byref(n64)),
        "edx_get_derived_state": (None, f16, buf, buf), "edx_set_render_target": (None, None, None), "edx_set_frame_sink": (None, None, None), "edx_set_frame_sink_signal": (None, None), "edx_flush_frame_sink": (None,), "edx_enable_peer_access": (None, 0), "edx_device_alloc": (None, 16, C.byref(out)),
        "edx_device_free": (None, None), "edx_read_device": (None, None, None, 0),
        "edx_set_screen_partition": (None, 0, 1), "edx_set_stream": (None, None), "edx_timer_begin": (None,),
        "edx_timer_end": (None, C.byref(ms)), "edx_set_profiling": (None, 1), "edx_get_stats": (None, C.byref(st)),
        "edx_set_option": (None, b"hiz", 1), "edx_debug_tile_residency": (None, C.cast(C.byref(u), C.POINTER(C.c_int))),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc < 0, (name, rc)
    assert not lib.edx_get_back_buffer(None)
    assert not lib.edx_device_color(None) and not lib.edx_device_depth(None)
    assert lib.edx_last_launch_count(None) == 0
    assert lib.edx_last_launch_list(None) == b""
    assert lib.edx_mesh_destroy(None, None) == 0          # destroying nothing is not an error
    lib.edx_destroy(None)
    lib.edx_last_error(None)
    untested = set(_lib.SYMBOLS) - set(calls) - {"edx_get_back_buffer", "edx_device_color", "edx_device_depth", "edx_last_launch_count", "edx_last_launch_list", "edx_device_count",
                                                 "edx_mesh_destroy", "edx_destroy", "edx_last_error", "edx_version", "edx_create"}
    assert not untested, untested
