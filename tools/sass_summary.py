#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (static instruction counts from `cuobjdump -sass`):
python tools/sass_summary.py > profiles/r2_sass_summary.txt.  The columns are the mnemonics that show how a kernel moves and
computes its data: UTMALDG / UBLKCP (TMA tensor / bulk copies), SYNCS (mbarrier), LDG / LDS / STG, LOP3 / SHF / PRMT (ALU pipe),
IDP (DP4A / DP2A), VIMNMX3, POPC, IMAD, BAR."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hyslam_b200", "lib", "libhyorb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTMALDG", "UBLKCP", "SYNCS", "LDG", "LDS", "STG", "STS", "LOP3", "SHF", "PRMT", "IDP", "VIMNMX3", "VIMNMX", "POPC", "IMAD", "BAR", "SHFL", "ATOM"]
kern, counts = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("hyorb::", "").replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for c in cols:
            if op == c or (c in ("ATOM",) and op.startswith("ATOM")) or (c == "BAR" and op == "BAR"):
                counts[kern][c] += 1
print("static SASS instruction counts per kernel, %s" % os.path.relpath(lib, ROOT))
print("%-34s %6s " % ("kernel", "total") + " ".join("%7s" % c for c in cols))
for k in sorted(counts):
    print("%-34s %6d " % (k[:34], counts[k]["_total"]) + " ".join("%7d" % counts[k][c] for c in cols))
