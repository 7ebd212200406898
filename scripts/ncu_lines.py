"""Aggregate an ncu report's per-instruction samples by CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep [kernel-regex] [top N]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kern: cmd += ["-k", "regex:" + kern]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict()
fname = func = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":                      # a source line row
        cur = (func, fname, r[0], r[1].strip()[:100])
        agg.setdefault(cur, [0, 0])
    if len(r) > 7 and r[2] != "":       # a SASS row under the current source line
        try:
            agg[cur][0] += int(r[6]); agg[cur][1] += int(r[7])
        except Exception: pass
byfunc = collections.defaultdict(list)
for k, v in agg.items(): byfunc[k[0]].append((v[0], v[1], k))
for f, items in byfunc.items():
    ts = sum(i[0] for i in items) or 1; ti = sum(i[1] for i in items) or 1
    print("==", f, "samples", ts, "warp-instructions", ti)
    for s, i, k in sorted(items, reverse=True)[:top]:
        print("%6.2f%% smp %6.2f%% inst  %s:%s | %s" % (100.0 * s / ts, 100.0 * i / ti, k[1], k[2], k[3]))
