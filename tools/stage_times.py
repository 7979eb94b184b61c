#!/usr/bin/env python
"""Per-stage device times of the compress pipeline for several corpora (quick A/B runs between kernel changes):
    python tools/stage_times.py [MiB] [family ...]      -> one JSON line per family"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from my_compress_b200 import api, datagen

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fams = sys.argv[2:] or ["uniform", "zipf", "phrases", "mixed"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for fam in fams:
    n = mib << 20
    x = torch.from_numpy(datagen.gen(fam, n, seed=20261018)).cuda()
    dc = api.DeviceCompressor(0, max_batch_blocks=min(1024, mib))
    out = torch.empty(api.compress_bound(n), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dc.compress_async(x, out)
    k = dc.finish()
    dc.enable_timing(True)
    reps = 5
    for _ in range(reps):
        flush.fill_(1)
        dc.compress_async(x, out)
        dc.finish()
    st = {a: round(b / reps, 4) for a, b in dc.stage_times().items()}
    tot = sum(st.values())
    print(json.dumps({"family": fam, "MiB": mib, "GBps": round(n / tot / 1e6, 2), "total_ms": round(tot, 3), "out": k, **st}))
    dc.close()
    del x, out
