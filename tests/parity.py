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
