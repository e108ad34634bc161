#!/usr/bin/env python
"""Top stall / instruction SASS lines of kernel #idx in an .ncu-rep.
    python tools/ncu_sass_top.py report.ncu-rep [kernel_index] [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split per kernel: header rows start with "Kernel Name"
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[kidx]:starts[kidx + 1]]
print(blk[0][1][:120])
h = blk[1]
ci, wi = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
data = []
for n, r in enumerate(blk[2:]):
    if len(r) > ci and r[0].startswith("0x"):
        data.append((n, int(r[ci] or 0), int(r[wi] or 0), r[1].strip()))
ti = sum(d[1] for d in data) or 1; tw = sum(d[2] for d in data) or 1
print("instr", ti, "samples", tw, "sass lines", len(data))
for n, i, w, s in sorted(data, key=lambda d: -d[2])[:topn]:
    print("%5d  stall %5.1f%%  inst %5.2f%%  %s" % (n, 100.0 * w / tw, 100.0 * i / ti, s[:100]))
