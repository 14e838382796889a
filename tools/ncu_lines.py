#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report (captured with --import-source on, built with -lineinfo):
    python tools/ncu_lines.py <rep> <kernel> [min_frac] [launch_skip]"""
import csv, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, fname, lines = None, "", []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":     # a source line row (SASS rows carry an address)
        lines.append((fname, r))
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for _, r in lines)
print("kernel", kern, "total samples", tot, "total warp instructions", sum(int(r[ii]) for _, r in lines))
for f, r in lines:
    if int(r[si]) >= max(1, tot * minf):
        print(f"{f:>16}:{r[0]:<5} samples={r[si]:>6} ({100*int(r[si])/tot:4.1f}%) inst={r[ii]:>8}  {r[1].strip()[:120]}")
