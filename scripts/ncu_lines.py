#!/usr/bin/env python
"""Per-source-line totals of an ncu report's source page (needs -lineinfo + --import-source on):
   ncu_lines.py <rep> [top_n]  ->  warp instructions executed, thread instructions, stall samples per file:line."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = collections.OrderedDict(); tot_i = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or not r[0]: continue
    ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    try: i, t, s = int(r[ie]), int(r[te]), int(r[ss])
    except ValueError: continue
    agg[(cur, int(r[0]), r[1].strip()[:90])] = (i, t, s); tot_i += i; tot_s += s
print("total warp instr %d, samples %d" % (tot_i, tot_s))
for (f, ln, src), (i, t, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% samples  lanes %4.1f  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), t / max(i, 1), f, ln, src))
