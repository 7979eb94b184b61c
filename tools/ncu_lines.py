#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report:
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]
Uses `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo builds and --import-source on)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# layout: blocks of [cuda line row] followed by its sass rows; find the header row of metrics
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])      # line -> [inst, thread inst, samples, text]
cur = None
for r in rows:
    if not r:
        continue
    if r[0] in ("Address", "Line No") or "Instructions Executed" in r:
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        i_smp = hdr.index("# Samples")
    except ValueError:
        continue
    key = r[0]
    if re.fullmatch(r"\d+", key):             # a CUDA source line row
        cur = int(key)
        agg[cur][3] = r[1]
        continue
    if cur is None or len(r) <= max(i_inst, i_thr, i_smp):
        continue
    try:
        agg[cur][0] += int(float(r[i_inst] or 0))
        agg[cur][1] += int(float(r[i_thr] or 0))
        agg[cur][2] += int(float(r[i_smp] or 0))
    except ValueError:
        pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for ln, v in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    lanes = v[1] / v[0] if v[0] else 0
    print("%5d  inst %5.1f%%  smp %5.1f%%  lanes %4.1f  | %s" % (ln, 100.0 * v[0] / tot_i, 100.0 * v[2] / tot_s, lanes, v[3].strip()[:110]))
