"""The C++ host API (include/edxraster/Renderer.h) through the headless viewer example: the program
mirrors RealtimeViewer/Main.cpp call for call; its frame must match the oracle on the same inputs."""
import os
import struct
import subprocess

import numpy as np
import pytest

from edxraster_b200 import scenes
from oracle import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


CUBE_OBJ = """# unit cube, quads, no normals
v -1 -1 -1
v 1 -1 -1
v 1 1 -1
v -1 1 -1
v -1 -1 1
v 1 -1 1
v 1 1 1
v -1 1 1
f 1 4 3 2
f 5 6 7 8
f 1 2 6 5
f 2 3 7 6
f 3 4 8 7
f 4 1 5 8
"""


def run_viewer(tmp_path, extra, msaa=0):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")])
    bmp, dump = str(tmp_path / "frame.bmp"), str(tmp_path / "dump.bin")
    out = subprocess.run([os.path.join(ROOT, "examples", "headless_viewer"), "5", bmp, dump] + extra, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return out, bmp, dump


def compare_with_oracle(bmp, dump, msaa=0):
    raw = open(dump, "rb").read()
    w, h, nv, nt = struct.unpack("4I", raw[:16])
    mats = np.frombuffer(raw, np.float32, 48, 16).reshape(3, 4, 4)
    verts = np.frombuffer(raw, np.float32, nv * 8, 16 + 192).reshape(nv, 8)
    idx = np.frombuffer(raw, np.uint32, nt * 3, 16 + 192 + nv * 32).reshape(nt, 3)
    o = orc.Oracle(w, h, 0)
    if msaa:
        o.set_msaa(msaa)
    o.set_transform(mats[0], mats[1], mats[2])
    o.set_shader(scenes.SHADER_BLINN_PHONG)
    o.render(verts, idx)
    ref = o.color()
    data = open(bmp, "rb").read()
    assert data[:2] == b"BM"
    off = struct.unpack("<I", data[10:14])[0]
    bw, bh = struct.unpack("<ii", data[18:26])
    assert (bw, bh) == (w, h)
    pix = np.frombuffer(data, np.uint8, w * h * 3, off).reshape(h, w, 3)[..., ::-1]
    assert np.abs(pix.astype(np.int32) - ref[..., :3].astype(np.int32)).max() <= 1
    return ref, nt


def test_obj_mesh_with_msaa_through_the_cpp_api(tmp_path):
    obj = tmp_path / "cube.obj"
    obj.write_text(CUBE_OBJ)
    out, bmp, dump = run_viewer(tmp_path, [str(obj), "2"])
    assert "Triangle Count: 12" in out.stdout               # 6 quads fanned into 12 triangles
    ref, nt = compare_with_oracle(bmp, dump, msaa=2)
    assert nt == 12
    assert int((ref[..., 3] > 0).sum()) > 20000
    assert len(np.unique(ref[..., 3])) > 2                   # anti-aliased silhouette: partial alpha values exist


def test_headless_viewer_matches_oracle(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")])
    bmp, dump = str(tmp_path / "frame.bmp"), str(tmp_path / "dump.bin")
    out = subprocess.run([os.path.join(ROOT, "examples", "headless_viewer"), "5", bmp, dump], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Triangle Count: 8192" in out.stdout            # LoadSphere default 64 x 64 (Mesh.h:43-44)
    raw = open(dump, "rb").read()
    w, h, nv, nt = struct.unpack("4I", raw[:16])
    mats = np.frombuffer(raw, np.float32, 48, 16).reshape(3, 4, 4)
    verts = np.frombuffer(raw, np.float32, nv * 8, 16 + 192).reshape(nv, 8)
    idx = np.frombuffer(raw, np.uint32, nt * 3, 16 + 192 + nv * 32).reshape(nt, 3)
    o = orc.Oracle(w, h, 0)
    o.set_transform(mats[0], mats[1], mats[2])
    o.set_shader(scenes.SHADER_BLINN_PHONG)
    o.render(verts, idx)
    ref = o.color()                                          # bottom-up RGBA
    data = open(bmp, "rb").read()
    assert data[:2] == b"BM"
    off = struct.unpack("<I", data[10:14])[0]
    bw, bh = struct.unpack("<ii", data[18:26])
    assert (bw, bh) == (w, h)
    pix = np.frombuffer(data, np.uint8, w * h * 3, off).reshape(h, w, 3)[..., ::-1]     # BGR -> RGB, bottom-up like ours
    d = np.abs(pix.astype(np.int32) - ref[..., :3].astype(np.int32))
    assert d.max() <= 1
    assert int((ref[..., 3] == 255).sum()) > 50000           # the sphere really is on screen


def test_frame_ring_through_the_cpp_api(tmp_path):
    # three frames in flight over a shared mesh: every frame equals the single-frame path, which equals the oracle
    out, bmp, dump = run_viewer(tmp_path, ["", "0", "3"])
    assert "3 frames in flight" in out.stdout
    assert "ring frames identical to the single-frame path: yes" in out.stdout
    compare_with_oracle(bmp, dump)


def test_textured_default_shader_through_the_cpp_api(tmp_path):
    # LoadSphere installs the constant 0.9 texture (Mesh.cpp:47); the viewer adds a 64x32 image and alternates the two
    # slots every five triangles; LambertianAlbedo + anisotropic 4x (Main.cpp:108)
    out, bmp, dump = run_viewer(tmp_path, ["", "0", "1", "3"])
    raw = open(dump, "rb").read()
    w, h, nv, nt = struct.unpack("4I", raw[:16])
    mats = np.frombuffer(raw, np.float32, 48, 16).reshape(3, 4, 4)
    verts = np.frombuffer(raw, np.float32, nv * 8, 16 + 192).reshape(nv, 8)
    idx = np.frombuffer(raw, np.uint32, nt * 3, 16 + 192 + nv * 32).reshape(nt, 3)
    ys, xs = np.mgrid[0:32, 0:64]
    tex = np.stack([(xs * 37 + ys * 11) & 255, np.where((xs // 4 + ys // 4) & 1, 230, 40), (xs * ys * 3) & 255, np.full_like(xs, 255)], axis=-1).astype(np.uint8)
    o = orc.Oracle(w, h, 0)
    o.set_transform(mats[0], mats[1], mats[2])
    o.set_shader(scenes.SHADER_LAMBERT_ALBEDO)
    o.set_textures([("constant", (0.9, 0.9, 0.9)), ("image", tex)], ((np.arange(nt) // 5) % 2).astype(np.uint32))
    o.set_texture_filter(3)
    o.render(verts, idx)
    ref = o.color()
    data = open(bmp, "rb").read()
    off = struct.unpack("<I", data[10:14])[0]
    pix = np.frombuffer(data, np.uint8, w * h * 3, off).reshape(h, w, 3)[..., ::-1]
    assert np.abs(pix.astype(np.int32) - ref[..., :3].astype(np.int32)).max() <= 1
    assert len(np.unique(pix.reshape(-1, 3), axis=0)) > 2000         # the image really is on the sphere
