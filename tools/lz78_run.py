#!/usr/bin/env python
"""FCX8 (lz78) throughput through the host-buffer calls:  python tools/lz78_run.py mixed 256"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from my_compress_b200 import api, datagen
family, mib = sys.argv[1], int(sys.argv[2])
n = mib << 20
data = datagen.gen(family, n, seed=20261018)
blob = api.compress_lz78(data)            # warm-up (allocations, module load)
t0 = time.perf_counter(); blob = api.compress_lz78(data); tc = time.perf_counter() - t0
back = api.decompress(blob)
t0 = time.perf_counter(); back = api.decompress(blob); td = time.perf_counter() - t0
# the reference's decoder drops a trailing 0x00 of EVERY block (mc:3701-3703): that is the expected output
blocks = [data[i:i + (1 << 20)] for i in range(0, n, 1 << 20)]
want = b"".join((blk[:-1] if blk[-1] == 0 else blk).tobytes() for blk in blocks)
ok = back == want
print("%-8s %5d MiB  FCX8 %d bytes (%.1f %%)  compress %.1f ms = %.3f GB/s   decompress %.1f ms = %.3f GB/s   round trip %s" % (
    family, mib, len(blob), 100.0 * len(blob) / n, tc * 1e3, n / tc / 1e9, td * 1e3, n / td / 1e9, "ok (reference decoder behaviour)" if ok else "MISMATCH"))
