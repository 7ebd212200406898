"""Markdown tables for DESIGN.md section 6 from the bench lines under gpurun_out/ (bench_n1.json, bench_ref.json, bench_n2/4/8.json).
usage: python scripts/design_table.py > /tmp/tables.md"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    p = os.path.join(ROOT, "gpurun_out", name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


d = load("bench_n1.json")
ref = load("bench_ref.json")
rf = d["roofline"]
print("**Headline (C2, one B200, `python bench.py`, SM clock %s MHz, no throttle reasons):**\n" % d["clocks"]["sm_mhz"])
print("| | value |")
print("|---|---|")
print("| frame, three in flight (`value`) | %.1f us/step = **%.0f Mtris/s**, %.1f Gpix/s |" % (d["ms_per_step"] * 1e3, d["value"], d["gpix_per_s"]))
print("| frame, one at a time | %.1f us = %.0f Mtris/s |" % (d["one_frame_in_flight"]["ms_per_step"] * 1e3, d["one_frame_in_flight"]["value"]))
print("| kernels per frame | %s |" % ", ".join("%s" % k for k in d["kernels_per_step"]))
print("| stage times, one frame profiled (CUDA events) | geometry %.1f, clip %.1f, resolve %.1f us |" % (rf["stage_ms"]["geom"] * 1e3, rf["stage_ms"]["clip"] * 1e3, rf["stage_ms"]["tile"] * 1e3))
print("| roofline, dominant kernel (`%s`) | %.0f GB/s of %.1f = **%.3f**; DRAM traffic %.1f MB for %.1f MB algorithmic |" % (
    rf["kernel"], rf["achieved"], rf["peak"], rf["frac"], (rf["traffic"] or 0) / 1e6, rf["algorithmic_bytes_per_launch"] / 1e6))
print("| roofline, whole frame | %.3f (frame-buffer bytes only: %.3f) |" % (rf["whole_frame_frac"], rf["fb_only_frac"]))
print("| end to end through the C ABI, host buffers | %.0f Mtris/s (%.2f ms/step: %.0f MB up, %.1f MB down per step - the PCIe floor); mesh resident: %.0f Mtris/s |" % (
    d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6, d["e2e"]["resident_mesh_value"]))
cb = d["cpu_baseline"]
print("| reference on the host (`oracle/_ref`, %d cores) | %.1f Mtris/s (%s) |" % (cb["cores"], cb["value"], cb["sample"].split(" after")[0]))
if ref:
    print("| `bench.py --impl reference` | %.1f Mtris/s, %.1f ms/step |" % (ref["value"], ref["ms_per_step"]))
print()
print("| config | frame, one at a time | three in flight | stage ms (geometry / clip+mid / tile+shade) | whole-frame HBM fraction | reference on %d host cores |" % cb["cores"])
print("|---|---|---|---|---|---|")
for k in ("C1", "C3", "C4"):
    o = d["other_configs"].get(k)
    if not o or "ms_per_frame" not in o:
        continue
    c = o.get("cpu_baseline", {})
    print("| %s | %.1f us (%.0f Mtris/s, %.1f Gpix/s) | %.1f us | %.1f / %.1f / %.1f us | %.3f | %s Mtris/s |" % (
        k, o["ms_per_frame"] * 1e3, o["mtris_per_s"], o["gpix_per_s"], o["ms_per_frame_3_in_flight"] * 1e3,
        o["stage_ms"]["geom"] * 1e3, o["stage_ms"]["clip"] * 1e3, o["stage_ms"]["tile"] * 1e3, o["hbm_frac_whole_frame"],
        ("%.4g" % c["value"]) if c.get("value") else "-"))
m1 = d["other_configs"].get("M1")
if m1 and "ms_per_frame" in m1:
    old = m1["shared_list_round1_routing"]
    print("| M1 (stress, not a BASELINE config) | %.0f us | %.0f us | %.0f / %.0f / %.0f us | | round 1's design (one shared list, no mid path): %.0f us / %.0f us |" % (
        m1["ms_per_frame"] * 1e3, m1["ms_per_frame_3_in_flight"] * 1e3, m1["stage_ms"]["geom"] * 1e3, m1["stage_ms"]["clip"] * 1e3,
        m1["stage_ms"]["tile"] * 1e3, old["ms_per_frame"] * 1e3, old["ms_per_frame_3_in_flight"] * 1e3))
print()
print("| GPUs | C2 (weak: one frame per GPU per step, every frame gathered to rank 0) | C5 (strong: 256 views of the 10M-triangle mesh, gathered to rank 0) |")
print("|---|---|---|")
base = d["value"]
c5base = d["other_configs"]["C5"]["ms_total"]
for n, name in ((1, "bench_n1.json"), (2, "bench_n2.json"), (4, "bench_n4.json"), (8, "bench_n8.json")):
    x = load(name)
    if not x or "C5" not in x.get("other_configs", {}):
        continue
    c5 = x["other_configs"]["C5"]
    nv = (x["roofline"].get("nvlink") or {})
    print("| %d | %.1f us/step, %.0f Mtris/s (x%.2f)%s | %.2f ms, %.0f views/s (x%.2f) |" % (
        n, x["ms_per_step"] * 1e3, x["value"], x["value"] / base,
        (", NVLink into rank 0: %.0f GB/s" % nv["achieved_gbs"]) if nv.get("achieved_gbs") else "",
        c5["ms_total"], c5["frames_per_s"], c5base / c5["ms_total"]))
