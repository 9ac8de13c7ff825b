#!/usr/bin/env python3
"""Dump the SASS of an .ncu-rep with executed counts and stall samples, in address order (read here, no GPU)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
data = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {n: i for i, n in enumerate(h)}
for r in data:
    print("%s %10s %6s %8s  %s" % (r[ci['Address']][-5:], r[ci['Instructions Executed']], r[ci['# Samples']],
                                 r[ci['Thread Instructions Executed']] if 'Thread Instructions Executed' in ci else '', r[ci['Source']]))
