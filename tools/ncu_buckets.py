#!/usr/bin/env python
"""Instruction / sample totals per bucket of SASS lines: python tools/ncu_buckets.py <rep> <kernel-regex> [bucket]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
bk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
isrc, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[iss]) for r in data)
for b in range(0, len(data), bk):
    ch = data[b:b + bk]
    e = sum(int(r[ie]) for r in ch); s = sum(int(r[iss]) for r in ch)
    mx = max(int(r[ie]) for r in ch)
    print(f"sass {b:5d}-{b+len(ch)-1:5d} inst%={e/tot*100:5.1f} samp%={s/tots*100:5.1f} maxexec={mx}")
