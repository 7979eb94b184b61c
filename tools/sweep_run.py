#!/usr/bin/env python
"""Stage times of one library variant:  FCX_LIBFCX=path python tools/sweep_run.py zipf 64 [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from my_compress_b200 import api, datagen

family = sys.argv[1]
mib = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = mib << 20
x = torch.from_numpy(datagen.gen(family, n, seed=20261018)).cuda()
dc = api.DeviceCompressor(0, max_batch_blocks=min(1024, mib))
out = torch.empty(api.compress_bound(n), dtype=torch.uint8, device="cuda")
for _ in range(3):
    dc.compress_async(x, out)
    k = dc.finish()
import hashlib
digest = hashlib.sha1(out[:k].cpu().numpy().tobytes()).hexdigest()[:12]
dc.enable_timing(True)
for _ in range(reps):
    dc.compress_async(x, out)
    dc.finish()
st = {k: v / reps for k, v in dc.stage_times().items()}
print("%-28s %-8s %5d MiB  match %8.3f  parse %7.3f  encode %6.3f  total %8.3f ms  sha1 %s" % (
    os.path.basename(os.environ.get("FCX_LIBFCX", "default")), family, mib, st["match"], st["parse"], st["encode"], sum(st.values()), digest))
