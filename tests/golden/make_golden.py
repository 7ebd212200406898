"""Generates tests/golden/golden.json and golden_frames.npz FROM THE REFERENCE ITSELF.

The reference ships no golden vectors, so these are made by running its own sources —
compiled unmodified against the EDXUtil stand-in, oracle/_ref (recipe: oracle/Makefile) — on reduced versions of
BASELINE.json's configs. They travel to boxes where /root/reference does not exist: the CPU suite checks that the
restatement (oracle/edx_oracle.cpp) reproduces them, the GPU suite checks that the CUDA path does, and where the
reference build is present it is checked against them as well.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

Hashes: per-pixel depth bits, per-pixel owner as the ORDINAL of the owning triangle in the set-up list (the only
identity the reference's records carry), RGBA8 back buffer, clip-space vertices, set-up records (six 28.4 integers;
z0 z1 z2 invW0 invW1 invW2 invDet).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from edxraster_b200 import scenes  # noqa: E402


def cases():
    """Reduced versions of BASELINE.json's configs; name -> scene. Sizes keep W, H even with W, H mod 32 in {0, 16..31}
    (the reference writes out of bounds otherwise, Rasterizer.h:92-95)."""
    return {
        "C1_small": scenes.config1(width=320, height=176, slices=32, stacks=32),
        "C2_small": scenes.config2(width=480, height=272, num_tris=20000),
        "C3_small": scenes.config3(width=384, height=208, num_tris=64),
        "C4_small": scenes.config4(width=480, height=272, quads_x=120, quads_z=96),
        # textured LambertianAlbedo shader: trilinear, anisotropic 16x, nearest + slots
        "TEX_plane_trilinear": scenes.textured_plane(width=320, height=176, tex_filter=2),
        "TEX_plane_aniso16": scenes.textured_plane(width=320, height=176, tex_filter=5),
        "TEX_sphere_nearest": scenes.textured_sphere(width=320, height=176, slices=32, stacks=32, tex_filter=0),
    }


def msaa_cases():
    """name -> (scene, SetMSAAMode level)"""
    c = cases()
    return {"C1_small_4x": (c["C1_small"], 2), "C4_small_8x": (c["C4_small"], 3), "C3_small_16x": (c["C3_small"], 4),
            "C1_small_32x": (c["C1_small"], 5)}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import parity
    assert parity.reference_available(), "oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists"
    out, frames = {}, {}
    for name, sc in cases().items():
        ref = parity.render_reference(sc, threads=2)
        ints, flts = ref["tris"]
        out[name] = {
            "source": "oracle/_ref (reference sources, compiled)",
            "width": sc.width, "height": sc.height, "triangles": sc.num_tris, "shader": int(sc.shader),
            "depth_sha256": digest(ref["depth"]), "winner_ordinal_sha256": digest(ref["winner_ord"]),
            "color_sha256": digest(ref["color"]) if ref["color_comparable"] else None,
            "clip_sha256": digest(ref["clip"]), "raster_tri_ints_sha256": digest(ints), "raster_tri_floats_sha256": digest(flts),
            "raster_tris": int(ints.shape[0]), "covered_pixels": int((ref["winner_ord"] != 0xFFFFFFFF).sum()),
            "fragments": int(ref["fragments"]),
        }
        if name == "C1_small":
            frames["C1_small_depth"] = ref["depth"]
            frames["C1_small_winner_ord"] = ref["winner_ord"]
            frames["C1_small_color"] = ref["color"]
        if name.startswith("TEX_"):
            frames[name + "_color"] = ref["color"]
    for name, (sc, level) in msaa_cases().items():
        ref = parity.render_reference(sc, threads=2, msaa=level)
        out[name] = {
            "source": "oracle/_ref (reference sources, compiled)",
            "width": sc.width, "height": sc.height, "triangles": sc.num_tris, "shader": int(sc.shader), "msaa_level": level,
            "color_sha256": digest(ref["color"]),
            "sample_depth_sha256": digest(np.stack([d for d, _ in ref["samples"]])),
            "sample_winner_ordinal_sha256": digest(np.stack([w for _, w in ref["samples"]])),
        }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **frames)
    print(json.dumps({k: v.get("covered_pixels") for k, v in out.items()}))


if __name__ == "__main__":
    main()
