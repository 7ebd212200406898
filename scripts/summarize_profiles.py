"""Turn gpurun_out/*.csv / *.ncu-rep into the tracked summaries under profiles/ (round-tagged)."""
import csv, glob, io, json, os, re, subprocess, sys


def kname(n):
    """"void tile_kernel<0>(FrameParams)" -> "tile_kernel"; the <1> instantiation (MSAA) keeps its suffix"""
    n = re.sub(r"^void\s+", "", n).split("(")[0]
    n = n.replace("edx::", "")
    base = re.sub(r"<.*>", "", n)
    if base == "tile_kernel" and ("<1>" in n or "<(bool)1>" in n):
        return "tile_kernel_msaa"
    return base

FRAME_KERNELS = ("geom_kernel", "cull_kernel", "vertex_kernel", "geom_list_kernel", "clip_kernel", "mid_kernel", "sort_big_kernel",
                 "bin_keys_kernel", "bin_fill_kernel", "bin_scan_kernel", "lean_resolve_kernel",
                 "tile_kernel", "shade_kernel", "msaa_resolve_kernel", "frame_end_kernel")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
lines = ["# ncu summaries (%s)" % tag, "",
         "Launch lists: `ncu --metrics gpu__time_duration.sum --clock-control none` over `scripts/profile_run.py --workload CX --frames 4`",
         "(cold-cache, serialised: compare SHARES, not absolutes). Full captures: `ncu --set full --clock-control none --import-source on`.", ""]
traffic = {}
for w in ("C1", "C2", "C3", "C4"):
    p = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % w)
    if not os.path.exists(p):
        continue
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    # frames end with frame_end_kernel; setup kernels (uploads, fills) precede the first geometry kernel
    frames, cur = [], []
    for r in rows[1:]:
        k = kname(r[ki])
        if k.split("_msaa")[0] not in FRAME_KERNELS:
            continue
        cur.append((k, float(r[vi].replace(",", ""))))
        if k == "frame_end_kernel":
            frames.append(cur); cur = []
    if not frames:
        continue
    steady = frames[1:] if len(frames) > 1 else frames          # the first frame sizes queues and launches every optional kernel
    names = [k for k, _ in steady[-1]]
    mean = {k: sum(v for f in steady for kk, v in f if kk == k) / max(1, sum(1 for f in steady for kk, _ in f if kk == k)) for k in names}
    tot = sum(mean.values()) or 1
    lines += ["## %s launch list (ns per launch, mean of frames 2-%d; frame 1 also launches: %s)" % (
                  w, len(frames), ", ".join(sorted(set(k for k, _ in frames[0]) - set(names))) or "nothing else"),
              "", "| kernel | ns | share of frame |", "|---|---|---|"]
    for k in names:
        lines.append("| %s | %.0f | %.1f %% |" % (k, mean[k], 100 * mean[k] / tot))
    lines.append("| (frame) | %.0f | |" % tot)
    lines.append("")
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "full_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    h = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
    idx = [(c, h.index(c)) for c in want if c in h]
    units = rows[1]
    lines += ["## full capture %s" % name, "", "| " + " | ".join(c for c, _ in idx) + " |", "|" + "---|" * len(idx)]
    last = {}
    for r in rows[2:]:                       # two frames are captured: keep each kernel's last launch
        last[kname(r[h.index("Kernel Name")])] = r
    for r in last.values():
        lines.append("| " + " | ".join((kname(r[i]) if c == "Kernel Name" else r[i] + " " + units[i]) for c, i in idx) + " |")
    for r in last.values():
        try:
            kn = kname(r[h.index("Kernel Name")])
            rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
            ur, uw = units[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_write.sum")]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic.setdefault(name[:2], {})[kn] = rd * mul.get(ur, 1) + wr * mul.get(uw, 1)
        except Exception:
            pass
    lines.append("")
open(os.path.join(out, "%s_ncu_summary.md" % tag), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1, sort_keys=True)
print("\n".join(lines[:60]))

# ---- launch list of the bench command itself ----
bl = os.path.join(ROOT, "gpurun_out", "bench_launches.csv")
if os.path.exists(bl):
    rows = [r for r in csv.reader(open(bl)) if len(r) > 10]
    h = rows[0]
    ki, vi, si = h.index("Kernel Name"), h.index("Metric Value"), h.index("Stream")
    md = ["# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-extra` (%s)" % tag, "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-extra`",
          "(the first 400 launches of the command: context set-up, the uploads of the four mesh copies, then the first warm-up",
          "frames of the headline loop with 3 frames in flight - three lanes = three streams. The first frame of a lane launches",
          "everything: sort_big_kernel, the four bin_* kernels and tile_kernel as well; once a frame has published its counters the",
          "idle ones are left out and a frame without large triangles ends in lean_resolve_kernel instead of tile_kernel, with",
          "mid_kernel as the catch-all: geom, clip, mid, lean_resolve, frame_end). Times are cold-cache and serialised by ncu:",
          "compare shares, not absolutes. Every kernel is ours; no library kernel runs in a step.", "",
          "| # | stream | kernel | ns |", "|---|---|---|---|"]
    tot = {}
    for n, r in enumerate(rows[1:]):
        k = kname(r[ki])
        v = float(r[vi].replace(",", ""))
        md.append("| %d | %s | %s | %.0f |" % (n, r[si], k, v))
        tot.setdefault(k, []).append(v)
    md += ["", "| kernel | launches | mean ns | share of the frame kernels |", "|---|---|---|---|"]
    steady = ("geom_kernel", "clip_kernel", "mid_kernel", "lean_resolve_kernel", "frame_end_kernel")      # a steady-state C2 frame
    fsum = sum(sum(v) / len(v) for k, v in tot.items() if k in steady) or 1
    for k, v in tot.items():
        share = "%.1f %%" % (100 * (sum(v) / len(v)) / fsum) if k in steady else ""
        md.append("| %s | %d | %.0f | %s |" % (k, len(v), sum(v) / len(v), share))
    open(os.path.join(out, "%s_bench_launches.md" % tag), "w").write("\n".join(md) + "\n")
