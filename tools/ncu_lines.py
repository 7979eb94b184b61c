#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report:
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top] [kernel-regex]
Uses `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo builds and --import-source on)."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_thr, i_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or not re.fullmatch(r"\d+", r[0]) or len(r) <= i_thr:
        continue
    try:
        v = [int(float(r[i_inst] or 0)), int(float(r[i_thr] or 0)), int(float(r[i_smp] or 0)), r[1]]
    except ValueError:
        continue
    ln = int(r[0])
    if ln in agg:
        agg[ln][0] += v[0]; agg[ln][1] += v[1]; agg[ln][2] += v[2]
    else:
        agg[ln] = v
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for ln, v in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    lanes = v[1] / v[0] if v[0] else 0
    print("%5d  inst %5.1f%%  smp %5.1f%%  lanes %4.1f  | %s" % (ln, 100.0 * v[0] / tot_i, 100.0 * v[2] / tot_s, lanes, v[3].strip()[:110]))
