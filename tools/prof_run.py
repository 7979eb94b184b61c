#!/usr/bin/env python
"""One compress pass of a synthetic workload, for ncu captures:  python tools/prof_run.py zipf 16"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from my_compress_b200 import api, datagen

family = sys.argv[1] if len(sys.argv) > 1 else "zipf"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = mib << 20
x = torch.from_numpy(datagen.gen(family, n, seed=20261018)).cuda()
dc = api.DeviceCompressor(0, max_batch_blocks=min(1024, mib))
out = torch.empty(api.compress_bound(n), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    dc.compress_async(x, out)
    k = dc.finish()
print(family, mib, "MiB ->", k, "bytes", dc.stats())
