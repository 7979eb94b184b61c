#!/usr/bin/env python
"""Huffman sections and block payloads of the library named by FCX_LIBFCX against the oracle (exit 1 on a
difference).  Used by tests/test_gpu_parity.py for the small-stage build, whose k_encode_w takes the direct
path (atomicOr from registers) for every sub-chunk denser than ~3 bits per symbol."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from _cases import stream_buffers
from _libs import oracle
from my_compress_b200 import api, datagen

o = oracle()
dc = api.DeviceCompressor(0, max_batch_blocks=4)
rng = np.random.default_rng(9)
bufs = list(stream_buffers())
bufs.append(("uniform300k", datagen.uniform_block(300000, 7, 0)))
bufs.append(("zipf200001", datagen.zipf_block(200001, 7, 0)))
bufs.append(("skew24577", np.minimum(rng.geometric(0.5, size=24577) - 1, 255).astype(np.uint8)))
for m in (9, 12, 18, 25):                                    # longest code m-1 bits: groups of 4 / 2 / 1 symbols
    fib = [1, 1]
    while len(fib) < m:
        fib.append(fib[-1] + fib[-2])
    base = rng.permutation(np.repeat(np.arange(m, dtype=np.uint8) * 7 + 3, fib))
    while len(base) < 8192 * 2 + 40:
        base = np.concatenate([base, rng.permutation(base)])
    bufs.append(("fib%d" % m, np.ascontiguousarray(base[: 8192 * 2 + 33])))
bad = 0
for name, s in bufs:
    t = torch.from_numpy(np.ascontiguousarray(s, dtype=np.uint8).copy()).cuda()
    pad = torch.zeros(((t.numel() + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")   # as tests/test_gpu_parity.py _dev
    pad[: t.numel()] = t
    if dc.probe_huffman_section(pad[: t.numel()]) != o.huffman_section(s):
        print("section differs:", name)
        bad += 1
for fam, n in (("uniform", 200000), ("zipf", 300001), ("phrases", 1 << 20)):
    x = datagen.gen_block(fam, n, 5, 0)
    if api.my_compress_file_lz77(x) != o.compress_block(x):
        print("payload differs:", fam, n)
        bad += 1
print("check_variant: %s, %d cases, %d differences" % (os.path.basename(os.environ.get("FCX_LIBFCX", "libfcx.so")), len(bufs) + 3, bad))
sys.exit(1 if bad else 0)
