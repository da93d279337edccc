#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel of an .ncu-rep captured with --import-source on.

    python tools/ncu_source_hotspots.py <report.ncu-rep> <kernel index, 1-based> [top N]
"""
import csv
import io
import os
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
data, hdr, col, fname, func = [], None, None, "?", "?"
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
    elif hdr and r[0].isdigit() and len(r) >= len(hdr):
        try:
            s = int(r[col["# Samples"]])
        except ValueError:
            continue
        data.append((s, fname, r))
print(func[:150])
tot = sum(s for s, _, _ in data) or 1
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for s, f, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[col[h]] or 0), h) for h in stalls), reverse=True)[:3]
    print(f"{100 * s / tot:5.1f}% {f[:14]:14s}:{r[0]:>4} {r[1].strip()[:95]:95s} | " + " ".join(f"{h[6:]}={v}" for v, h in st if v))
