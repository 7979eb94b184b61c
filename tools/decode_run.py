#!/usr/bin/env python
"""Decode throughput of the host-buffer call (pinned buffers):  python tools/decode_run.py zipf 64 [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from my_compress_b200 import api, datagen

family = sys.argv[1] if len(sys.argv) > 1 else "zipf"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = mib << 20
data = datagen.gen(family, n, seed=20261018)
blob = api.compress(data)
src = torch.empty(len(blob), dtype=torch.uint8).pin_memory()
src.numpy()[:] = np.frombuffer(blob, np.uint8)
dst = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
got = api.decompress_into(src.numpy(), dst.numpy())
assert got == n and bytes(dst.numpy()[:n]) == data.tobytes(), "round trip failed"
t0 = time.perf_counter()
for _ in range(reps):
    api.decompress_into(src.numpy(), dst.numpy())
dt = (time.perf_counter() - t0) / reps
print("%-8s %5d MiB  container %d bytes  decode %.1f ms  %.3f GB/s (output bytes; H2D + kernel + D2H, pinned host buffers)" % (
    family, mib, len(blob), dt * 1e3, n / dt / 1e9))
