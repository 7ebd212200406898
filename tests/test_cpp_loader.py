"""The C++ layer's OBJ / MTL / BMP loader (include/edxraster/Renderer.h, Mesh::LoadMesh - Utils/Mesh.cpp:11-34) is host
code: compile a small program against the header and check what it builds, no GPU needed."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "edxraster/Renderer.h"
#include <cstdio>
using namespace edx_b200;
int main(int argc, char** argv) {
    Mesh m;
    if (!m.LoadMesh(Vector3(1, 2, 3), Vector3(2, 2, 2), Vector3(0, 0, 0), argv[1])) { std::printf("load failed\n"); return 1; }
    std::printf("tris %u verts %u textures %zu\nids", m.GetIndexBuffer()->GetTriangleCount(), m.GetVertexBuffer()->GetVertexCount(), m.GetTextureCount());
    for (uint i : m.GetTextureIds()) std::printf(" %u", i);
    for (size_t k = 0; k < m.GetTextureCount(); k++) {
        int kind; float c[3]; uint w, h; const _byte* px;
        m.GetTexture(k, kind, c, w, h, px);
        std::printf("\nslot %zu kind %d color %.3f %.3f %.3f size %ux%u", k, kind, c[0], c[1], c[2], w, h);
        if (kind == 1) { std::printf(" texels"); for (uint i = 0; i < w * h * 4; i++) std::printf(" %d", px[i]); }
    }
    const float* v = (const float*)m.GetVertexBuffer()->GetBuffer();
    std::printf("\nv0 %.3f %.3f %.3f uv %.3f %.3f\n", v[0], v[1], v[2], v[6], v[7]);
    const BoundingBox b = m.GetBounds();
    std::printf("bounds %.3f %.3f %.3f .. %.3f %.3f %.3f\n", b.mMin.x, b.mMin.y, b.mMin.z, b.mMax.x, b.mMax.y, b.mMax.z);
    Scene scene; scene.AddMesh(new Mesh); std::printf("scene %zu\n", scene.GetMeshCount());
    return 0;
}
'''


def write_bmp(path, img, bpp=24, top_down=False):
    h, w = img.shape[:2]
    px = img[..., [2, 1, 0]] if bpp == 24 else np.concatenate([img[..., [2, 1, 0]], np.full((h, w, 1), 255, np.uint8)], axis=-1)
    stride = (w * (bpp // 8) + 3) & ~3
    order = range(h) if top_down else range(h - 1, -1, -1)
    rows = b"".join(px[y].tobytes() + b"\0" * (stride - w * (bpp // 8)) for y in order)
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, 0, len(rows), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + rows)


def build(tmp_path):
    src, exe = tmp_path / "loader.cpp", tmp_path / "loader"
    src.write_text(PROGRAM)
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L" + os.path.join(ROOT, "edxraster_b200"), "-ledxraster_b200", "-Wl,-rpath," + os.path.join(ROOT, "edxraster_b200")])
    return str(exe)


def test_obj_with_materials_and_bmp_textures(tmp_path):
    exe = build(tmp_path)
    rng = np.random.default_rng(4)
    a, b = rng.integers(0, 256, (3, 5, 3), dtype=np.uint8), rng.integers(0, 256, (2, 2, 3), dtype=np.uint8)
    write_bmp(str(tmp_path / "a.bmp"), a)                               # 24-bit, bottom-up, padded rows (5 * 3 = 15 -> 16)
    write_bmp(str(tmp_path / "b.bmp"), b, bpp=32, top_down=True)
    (tmp_path / "m.mtl").write_text("# comment\nnewmtl red\n  Kd 0.8 0.2 0.1\nnewmtl pic\nKd 1 1 1\nmap_Kd a.bmp\nnewmtl missing\nKd 0.1 0.2 0.3\nmap_Kd nope.bmp\nnewmtl pic32\nmap_Kd b.bmp\n")
    (tmp_path / "c.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0.25 0.5\nvt 1 0\nvt 1 1\nvt 0 1\n"
                                    "f 1/1 2/2 3/3\nusemtl pic\nf 1/1 2/2 3/3 4/4\nusemtl pic32\nf 1/1 3/3 4/4\nusemtl missing\nf -4/-4 -3/-3 -2/-2\nusemtl red\nf 1/1 2/2 4/4\n")
    out = subprocess.run([exe, str(tmp_path / "c.obj")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "tris 6 verts 4 textures 4"
    assert lines[1] == "ids 0 1 1 3 2 0"                               # before any usemtl: the first material
    assert lines[2].startswith("slot 0 kind 0 color 0.800 0.200 0.100")
    assert lines[3].startswith("slot 1 kind 1") and "size 5x3" in lines[3]
    got = np.array(lines[3].split("texels")[1].split(), np.uint8).reshape(3, 5, 4)
    np.testing.assert_array_equal(got[..., :3], a)                      # first row = top of the picture
    assert (got[..., 3] == 255).all()
    assert lines[4].startswith("slot 2 kind 0 color 0.100 0.200 0.300")  # unreadable map_Kd: the material's Kd
    got = np.array(lines[5].split("texels")[1].split(), np.uint8).reshape(2, 2, 4)
    np.testing.assert_array_equal(got[..., :3], b)
    assert lines[6] == "v0 1.000 2.000 3.000 uv 0.250 0.500"             # scale, then translation
    assert lines[7] == "bounds 1.000 2.000 3.000 .. 3.000 4.000 3.000"   # Mesh::GetBounds (Mesh.h:63-66)
    assert lines[8] == "scene 1"


def test_obj_without_materials_gets_the_constant_white_slot(tmp_path):
    exe = build(tmp_path)
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    out = subprocess.run([exe, str(tmp_path / "t.obj")], capture_output=True, text=True, timeout=60)
    lines = out.stdout.splitlines()
    assert lines[0] == "tris 1 verts 3 textures 1" and lines[1] == "ids 0"
    assert lines[2].startswith("slot 0 kind 0 color 0.900 0.900 0.900")   # Mesh.cpp:47,66
