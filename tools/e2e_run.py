#!/usr/bin/env python
"""e2e GB/s of the host-buffer call for one library variant: FCX_LIBFCX=... python tools/e2e_run.py zipf 64"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from my_compress_b200 import api, datagen
family, mib = sys.argv[1], int(sys.argv[2])
n = mib << 20
src = torch.from_numpy(datagen.gen(family, n, seed=20261018)).pin_memory().numpy()
dst = torch.empty(api.compress_bound(n), dtype=torch.uint8).pin_memory().numpy()
for _ in range(3): api.compress_into(src, dst, 1)
reps = 10 if mib <= 128 else 4
t0 = time.perf_counter()
for _ in range(reps): api.compress_into(src, dst, 1)
dt = (time.perf_counter() - t0) / reps
print("%-24s %-8s %5d MiB  e2e %7.2f GB/s  (%.3f ms)" % (os.path.basename(os.environ.get("FCX_LIBFCX", "default")), family, mib, n / dt / 1e9, dt * 1e3))
