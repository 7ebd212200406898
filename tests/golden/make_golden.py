"""Generates tests/golden/golden.json and golden_frames.npz from the CPU oracle.

The reference has no golden vectors of its own and cannot be built or imported here,
so these fixtures are produced by the oracle (oracle/edx_oracle.cpp, parity unpinned) after it passed
the hand-derived known-answer tests in tests/test_oracle_kats.py. They freeze the oracle's behaviour:
the CPU suite checks the oracle still reproduces them, the GPU suite checks the CUDA path does.

    python tests/golden/make_golden.py
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
    """Reduced versions of BASELINE.json's configs; name -> scene"""
    return {
        "C1_small": scenes.config1(width=320, height=180, slices=32, stacks=32),
        "C2_small": scenes.config2(width=480, height=270, num_tris=20000),
        "C3_small": scenes.config3(width=384, height=216, num_tris=64),
        "C4_small": scenes.config4(width=480, height=270, quads_x=120, quads_z=96),
        # textured LambertianAlbedo shader: trilinear, anisotropic 16x, nearest + slots
        "TEX_plane_trilinear": scenes.textured_plane(width=320, height=180, tex_filter=2),
        "TEX_plane_aniso16": scenes.textured_plane(width=320, height=180, tex_filter=5),
        "TEX_sphere_nearest": scenes.textured_sphere(width=320, height=180, slices=32, stacks=32, tex_filter=0),
    }


def msaa_cases():
    """name -> (scene, SetMSAAMode level)"""
    c = cases()
    return {"C1_small_4x": (c["C1_small"], 2), "C4_small_8x": (c["C4_small"], 3)}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import parity
    out, frames = {}, {}
    for name, sc in cases().items():
        ref = parity.render_oracle(sc, threads=2)
        ints, flts = ref["tris"]
        out[name] = {
            "width": sc.width, "height": sc.height, "triangles": sc.num_tris, "shader": int(sc.shader),
            "depth_sha256": digest(ref["depth"]), "winner_sha256": digest(ref["winner"]), "color_sha256": digest(ref["color"]),
            "clip_sha256": digest(ref["clip"]), "raster_tri_ints_sha256": digest(ints), "raster_tri_floats_sha256": digest(flts),
            "raster_tris": int(ints.shape[0]), "covered_pixels": int((ref["winner"] != 0xFFFFFFFF).sum()),
            "covered_samples": int(ref["stats"]["covered_samples"]),
        }
        if name == "C1_small":
            frames["C1_small_depth"] = ref["depth"]
            frames["C1_small_winner"] = ref["winner"]
            frames["C1_small_color"] = ref["color"]
        if name.startswith("TEX_"):
            frames[name + "_color"] = ref["color"]
    for name, (sc, level) in msaa_cases().items():
        ref = parity.render_oracle(sc, threads=2, msaa=level)
        out[name] = {
            "width": sc.width, "height": sc.height, "triangles": sc.num_tris, "shader": int(sc.shader), "msaa_level": level,
            "color_sha256": digest(ref["color"]),
            "sample_depth_sha256": digest(np.stack([d for d, _ in ref["samples"]])),
            "sample_winner_sha256": digest(np.stack([w for _, w in ref["samples"]])),
        }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **frames)
    print(json.dumps({k: v.get("covered_pixels") for k, v in out.items()}))


if __name__ == "__main__":
    main()
