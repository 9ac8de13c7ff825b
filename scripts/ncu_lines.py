#!/usr/bin/env python3
"""Per source line of an .ncu-rep captured with --import-source on (read here, no GPU): warp-instructions executed,
their share, stall samples, shared-memory wavefronts.  Usage: python scripts/ncu_lines.py x.ncu-rep [rows_per_launch]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
ci = {}
for i, n in enumerate(h):
    ci.setdefault(n, i)
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[2] != "-":
        continue
    try:
        lines.append((int(r[0]), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), int(r[ci["L1 Wavefronts Shared"]] or 0), r[1]))
    except ValueError:
        pass
tot = sum(l[1] for l in lines); ts = sum(l[2] for l in lines); tw = sum(l[3] for l in lines)
print("total warp-instructions %d, samples %d, shared wavefronts %d" % (tot, ts, tw))
for ln, ex, sm, wf, src in sorted(lines):
    if ex * 500 >= tot or sm * 500 >= ts:
        print("%5d %7.2f%% instr %6.2f%% samp %6.2f%% wf %s %s" % (ln, 100.0 * ex / tot, 100.0 * sm / ts, 100.0 * wf / max(tw, 1),
              ("%7.2f/row" % (ex / rows_n)) if rows_n else "", src.strip()[:110]))
