"""Shared parity harness: render one scene with the oracle and with the CUDA path, compare."""
import numpy as np

from edxraster_b200 import renderer as R
from oracle import orc


def render_oracle(scene, threads=0, shader=None, hierarchical=True, msaa=0):
    o = orc.Oracle(scene.width, scene.height, threads)
    if msaa:
        o.set_msaa(msaa)
    o.set_transform(scene.mv, scene.proj, scene.raster)
    o.set_shader(scene.shader if shader is None else shader)
    o.set_hierarchical(hierarchical)
    if scene.get("textures"):
        o.set_textures(scene["textures"], scene.get("tex_ids"))
        o.set_texture_filter(scene.get("tex_filter", 2))
    o.render(scene.vertices, scene.indices)
    out = {"color": o.color(), "depth": o.depth(), "winner": o.winner(), "clip": o.clip_verts(),
           "tris": o.raster_tris(), "stats": o.stats(), "derived": o.derived()}
    if msaa:
        out["samples"] = [(o.depth(k), o.winner(k)) for k in range(1 << msaa)]
    o.close()
    return out


def render_gpu(scene, shader=None, options=None, hierarchical=True, stages=True, renderer=None, msaa=0):
    r = renderer or R.Renderer(0)
    r.Initialize(scene.width, scene.height)
    r.SetMSAAMode(msaa)
    r.SetTransform(scene.mv, scene.proj, scene.raster)
    r.SetPixelShader(scene.shader if shader is None else shader)
    r.SetHierarchicalRasterize(hierarchical)
    r.SetCaptureIds(True)
    for k, v in (options or {}).items():
        r.SetOption(k, v)
    m = r.CreateMesh(scene.vertices, scene.indices)
    if scene.get("textures"):
        m.SetTextures(scene["textures"], scene.get("tex_ids"))
    r.SetTextureFilter(scene.get("tex_filter", 2))
    r.RenderMesh(m)
    out = {"color": r.GetBackBuffer().copy(), "depth": r.GetDepthBuffer(), "winner": r.GetWinnerIds(),
           "stats": r.GetStats(), "derived": r.DerivedState()}
    if msaa:
        out["samples"] = [r.GetSample(k) for k in range(1 << msaa)]
    if stages:
        out["clip"] = r.DebugClipVertices(m)
        out["tris"] = r.DebugRasterTriangles(m)
    m.Release()
    if renderer is None:
        r.close()
    return out


def compare(ref, got, color_tol=1):
    """Returns a dict of mismatch counts; all zeros == parity."""
    rep = {}
    rep["depth_bits"] = int((ref["depth"].view(np.uint32) != got["depth"].view(np.uint32)).sum())
    rep["winner"] = int((ref["winner"] != got["winner"]).sum())
    dc = np.abs(ref["color"].astype(np.int32) - got["color"].astype(np.int32))
    rep["color_gt_tol"] = int((dc > color_tol).any(axis=-1).sum())
    rep["color_max_diff"] = int(dc.max()) if dc.size else 0
    rep["color_exact_frac"] = float((dc == 0).all(axis=-1).mean()) if dc.size else 1.0
    if "samples" in ref and "samples" in got:
        rep["sample_depth_bits"] = sum(int((a[0].view(np.uint32) != b[0].view(np.uint32)).sum()) for a, b in zip(ref["samples"], got["samples"]))
        rep["sample_winner"] = sum(int((a[1] != b[1]).sum()) for a, b in zip(ref["samples"], got["samples"]))
    if "clip" in got:
        rep["clip_bits"] = int((ref["clip"].view(np.uint32) != got["clip"].view(np.uint32)).sum())
        ri, rf = ref["tris"]
        gi, gf = got["tris"]
        rep["tri_count"] = (int(ri.shape[0]), int(gi.shape[0]))
        if ri.shape == gi.shape:
            rep["tri_int_mismatch"] = int((ri != gi).sum())
            rep["tri_float_mismatch"] = int((rf.view(np.uint32) != gf.view(np.uint32)).sum())
        else:
            rep["tri_int_mismatch"] = rep["tri_float_mismatch"] = -1
    mvp_r, eye_r, l_r = ref["derived"]
    mvp_g, eye_g, l_g = got["derived"]
    rep["derived_bits"] = int((mvp_r.view(np.uint32) != mvp_g.view(np.uint32)).sum() + (eye_r.view(np.uint32) != eye_g.view(np.uint32)).sum()
                              + (l_r.view(np.uint32) != l_g.view(np.uint32)).sum())
    return rep


def is_parity(rep):
    ok = rep["depth_bits"] == 0 and rep["winner"] == 0 and rep["color_gt_tol"] == 0 and rep["derived_bits"] == 0
    ok = ok and rep.get("sample_depth_bits", 0) == 0 and rep.get("sample_winner", 0) == 0
    if "clip_bits" in rep:
        ok = ok and rep["clip_bits"] == 0 and rep["tri_int_mismatch"] == 0 and rep["tri_float_mismatch"] == 0
    return ok


# ---- the reference itself (oracle/_ref: its own sources compiled against the EDXUtil stand-in) ----

def reference_available():
    from oracle import ref
    return ref.available()


def to_ordinal(winner, tri_ints):
    """Owner ids (prim = submitted triangle * 8 + fan index) -> position of that triangle in the set-up list, the
    only identity the reference's own records carry. `tri_ints[:, 0]` is the prim column of the oracle / CUDA dump."""
    prims = np.asarray(tri_ints)[:, 0].astype(np.int64)
    out = np.full(winner.shape, 0xFFFFFFFF, np.uint32)
    m = winner != 0xFFFFFFFF
    if m.any():
        order = np.argsort(prims, kind="stable")
        pos = np.searchsorted(prims[order], winner[m].astype(np.int64))
        assert (prims[order][pos] == winner[m]).all(), "owner id without a set-up record"
        out[m] = order[pos].astype(np.uint32)
    return out


def render_reference(scene, threads=0, shader=None, hierarchical=True, msaa=0):
    """Render with the reference's own code. Shaders the reference can run: 1 Blinn-Phong, 3 LambertianAlbedo; for a
    depth-only (0) or plain Lambert (2) scene the default shader runs and only geometry / depth / owners are comparable."""
    from oracle import ref
    shader = scene.shader if shader is None else shader
    r = ref.Reference(scene.width, scene.height, threads)
    try:
        if msaa:
            r.set_msaa(msaa)
        r.set_transform(scene.mv, scene.proj, scene.raster)
        r.set_shader(shader if shader in (1, 3) else 3)
        r.set_hierarchical(hierarchical)
        r.set_mesh(scene.vertices, scene.indices, scene.get("textures"), scene.get("tex_ids"))
        r.set_texture_filter(scene.get("tex_filter", 2))
        r.render()
        out = {"color": r.color(), "depth": r.depth(), "winner_ord": r.winner_ordinal(), "clip": r.clip_verts(),
               "tris": r.raster_tris(), "derived": r.derived(), "fragments": r.num_fragments(), "color_comparable": shader in (1, 3),
               "threads": r.threads}
        if msaa:
            out["samples"] = [(r.depth(k), r.winner_ordinal(k)) for k in range(1 << msaa)]
    finally:
        r.close()
    return out


def compare_reference(ref, got):
    """`ref` from render_reference, `got` from render_oracle / render_gpu. All zeros == identical to the reference."""
    rep = {}
    rep["depth_bits"] = int((ref["depth"].view(np.uint32) != got["depth"].view(np.uint32)).sum())
    if "tris" in got:
        ri, rf = ref["tris"]
        gi, gf = got["tris"]
        rep["tri_count"] = (int(ri.shape[0]), int(gi.shape[0]))
        if ri.shape[0] == gi.shape[0]:
            rep["tri_int_mismatch"] = int((ri != gi[:, 1:]).sum())
            rep["tri_float_mismatch"] = int((rf.view(np.uint32) != gf.view(np.uint32)).sum())
            rep["winner"] = int((ref["winner_ord"] != to_ordinal(got["winner"], gi)).sum())
            if "samples" in ref and "samples" in got:
                rep["sample_depth_bits"] = sum(int((a[0].view(np.uint32) != b[0].view(np.uint32)).sum()) for a, b in zip(ref["samples"], got["samples"]))
                rep["sample_winner"] = sum(int((a[1] != to_ordinal(b[1], gi)).sum()) for a, b in zip(ref["samples"], got["samples"]))
        else:
            rep["tri_int_mismatch"] = rep["tri_float_mismatch"] = rep["winner"] = -1
    if "clip" in got:
        rep["clip_bits"] = int((ref["clip"].view(np.uint32) != got["clip"].view(np.uint32)).sum())
    if ref["color_comparable"]:
        dc = np.abs(ref["color"].astype(np.int32) - got["color"].astype(np.int32))
        rep["color_max_diff"] = int(dc.max()) if dc.size else 0
        rep["color_diff_pixels"] = int((dc > 0).any(axis=-1).sum())
    mvp_r, eye_r = ref["derived"]
    mvp_g, eye_g = got["derived"][0], got["derived"][1]
    rep["derived_bits"] = int((mvp_r.view(np.uint32) != mvp_g.view(np.uint32)).sum() + (eye_r.view(np.uint32) != eye_g.view(np.uint32)).sum())
    return rep


def is_reference_parity(rep, color_tol=0):
    keys = ("depth_bits", "tri_int_mismatch", "tri_float_mismatch", "winner", "clip_bits", "sample_depth_bits", "sample_winner", "derived_bits")
    return all(rep.get(k, 0) == 0 for k in keys) and rep.get("color_max_diff", 0) <= color_tol
