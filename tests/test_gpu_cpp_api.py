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


def _write_bmp24(path, img):
    """img: H x W x 3 uint8 RGB, row 0 = top. Classic bottom-up 24-bit BMP."""
    h, w = img.shape[:2]
    stride = (w * 3 + 3) & ~3
    rows = b"".join(img[y, :, ::-1].tobytes() + b"\0" * (stride - w * 3) for y in range(h - 1, -1, -1))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(rows), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + rows)


def read_dump(dump):
    raw = open(dump, "rb").read()
    w, h, nv, nt = struct.unpack("4I", raw[:16])
    mats = np.frombuffer(raw, np.float32, 48, 16).reshape(3, 4, 4)
    verts = np.frombuffer(raw, np.float32, nv * 8, 16 + 192).reshape(nv, 8)
    off = 16 + 192 + nv * 32
    idx = np.frombuffer(raw, np.uint32, nt * 3, off).reshape(nt, 3)
    off += nt * 12
    ntex = struct.unpack_from("I", raw, off)[0]; off += 4
    tex = []
    for _ in range(ntex):
        kind, r, g, b, tw, th = struct.unpack_from("i3f2I", raw, off); off += 24
        if kind == 1:
            tex.append(("image", np.frombuffer(raw, np.uint8, tw * th * 4, off).reshape(th, tw, 4))); off += tw * th * 4
        else:
            tex.append(("constant", (r, g, b)))
    nids = struct.unpack_from("I", raw, off)[0]; off += 4
    ids = np.frombuffer(raw, np.uint32, nids, off)
    return w, h, mats, verts, idx, tex, ids


def test_obj_materials_become_texture_slots(tmp_path):
    # Mesh::LoadMesh (Utils/Mesh.cpp:22-31): one slot per material - ImageTexture for map_Kd, ConstantTexture2D(Kd)
    # otherwise - and one slot id per face
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (8, 16, 3), dtype=np.uint8)
    _write_bmp24(str(tmp_path / "tex.bmp"), img)
    (tmp_path / "cube.mtl").write_text("newmtl red\nKd 0.8 0.2 0.1\nnewmtl img\nKd 1 1 1\nmap_Kd tex.bmp\n")
    verts = "".join("v %d %d %d\n" % p for p in [(1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)])
    vts = "vt 0 0\nvt 3 0\nvt 3 2\nvt 0 2\n"
    faces = ["1 4 3 2", "5 6 7 8", "1 2 7 6", "2 3 8 7", "3 4 5 8", "4 1 6 5"]
    body = "mtllib cube.mtl\n" + verts + vts
    for k, fc in enumerate(faces):
        body += "usemtl %s\n" % ("img" if k % 2 == 0 else "red")
        body += "f " + " ".join("%s/%d" % (v, t + 1) for t, v in enumerate(fc.split())) + "\n"
    obj = tmp_path / "cube.obj"
    obj.write_text(body)
    out, bmp, dump = run_viewer(tmp_path, [str(obj), "0", "1", "2", "mtl"])
    assert "Triangle Count: 12" in out.stdout
    w, h, mats, v, idx, tex, ids = read_dump(dump)
    assert [t[0] for t in tex] == ["constant", "image"]
    np.testing.assert_allclose(tex[0][1], (0.8, 0.2, 0.1), rtol=1e-6)
    np.testing.assert_array_equal(tex[1][1][..., :3], img)                 # row 0 = top of the picture
    assert (tex[1][1][..., 3] == 255).all()
    np.testing.assert_array_equal(ids, np.repeat([1, 0, 1, 0, 1, 0], 2))   # two triangles per quad, slot of its usemtl
    o = orc.Oracle(w, h, 0)
    o.set_transform(mats[0], mats[1], mats[2])
    o.set_shader(scenes.SHADER_LAMBERT_ALBEDO)
    o.set_textures(tex, ids)
    o.set_texture_filter(2)
    o.render(v, idx)
    ref = o.color()
    data = open(bmp, "rb").read()
    off = struct.unpack("<I", data[10:14])[0]
    pix = np.frombuffer(data, np.uint8, w * h * 3, off).reshape(h, w, 3)[..., ::-1]
    assert np.abs(pix.astype(np.int32) - ref[..., :3].astype(np.int32)).max() <= 1
    assert len(np.unique(pix.reshape(-1, 3), axis=0)) > 500


def test_frame_farm_through_the_cpp_api():
    # examples/farm_viewer: views dealt over every GPU of the box (one on a single-GPU box), three frames in flight per
    # GPU, each finished frame pushed into GPU 0's store by the copy engine (edx_set_frame_sink); the program compares
    # every farmed frame with the single-Renderer frame of the same view and exits 0 iff all are identical
    exe = os.path.join(ROOT, "examples", "farm_viewer")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")])
    out = subprocess.run([exe, "13", "0", "120"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 of 13 farmed frames differ" in out.stdout
