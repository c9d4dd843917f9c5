#!/usr/bin/env python
"""Top source lines of one kernel launch in an .ncu-rep (captured with --import-source on, built with -lineinfo):
warp instructions, active lanes, stall samples per line.
usage: python tools/ncu_hot_lines.py rep.ncu-rep [launch_skip] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path", "Function Name"):
        if r[0] != "Function Name": fname = r[1].split("/")[-1]
        elif not fname: print("#", r[1][:90])
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and r and r[0]:
        g = lambda k: int(r[hdr.index(k) - len(hdr)])       # from the end: source text with quotes can split into extra columns
        try:
            lines.append((fname, int(r[0]), r[1].strip(), g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples")))
        except ValueError:
            pass
ti = sum(l[3] for l in lines); ts = sum(l[5] for l in lines)
print(f"# total warp instructions {ti:,}  samples {ts:,}")
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'lanes':>5s}  source")
for l in sorted(lines, key=lambda l: -l[5])[:top]:
    print(f"{l[0] + ':' + str(l[1]):28s} {100 * l[3] / ti:6.2f} {100 * l[5] / max(ts, 1):6.2f} {l[4] / max(l[3], 1):5.1f}  {l[2][:110]}")
# per-file totals
agg = {}
for l in lines:
    a = agg.setdefault(l[0], [0, 0, 0]); a[0] += l[3]; a[1] += l[5]; a[2] += l[4]
print("\n# per file: inst% samp% lanes")
for f, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{f:32s} {100 * a[0] / ti:6.2f} {100 * a[1] / max(ts, 1):6.2f} {a[2] / max(a[0], 1):5.1f}")
if len(sys.argv) > 4:                      # line ranges "file:lo-hi,..." summed
    print("\n# ranges")
    for spec in sys.argv[4].split(","):
        f, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        sel = [l for l in lines if l[0] == f and lo <= l[1] <= hi]
        print(f"{spec:36s} inst {100 * sum(l[3] for l in sel) / ti:6.2f}%  samples {100 * sum(l[5] for l in sel) / max(ts, 1):6.2f}%")
