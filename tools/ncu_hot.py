#!/usr/bin/env python
"""Hot SASS regions of one kernel in an ncu report: python tools/ncu_hot.py <rep> <kernel-regex> [launch-skip]"""
import csv, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
isrc, ie, iss, it = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[iss]) for r in data); tt = sum(int(r[it]) for r in data)
print(rows[0][1][:80]); print("warp inst", tot, "avg active lanes %.1f" % (tt / tot), "samples", tots, "nsass", len(data))
groups = []
for i, r in enumerate(data):
    e, s, op = int(r[ie]), int(r[iss]), r[isrc].strip()
    op = op.split()[1] if op.startswith('@') else op.split()[0]
    if groups and groups[-1][1] == e: groups[-1][2] += 1; groups[-1][3] += s; groups[-1][4].append(op)
    else: groups.append([i, e, 1, s, [op]])
for g in groups:
    if g[1] * g[2] > tot * 0.01 or g[3] > tots * 0.02:
        print(f"sass#{g[0]:4d} exec={g[1]:9d} n={g[2]:3d} inst%={g[1]*g[2]/tot*100:5.1f} samp%={g[3]/tots*100:5.1f} {Counter(g[4]).most_common(7)}")
