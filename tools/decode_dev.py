#!/usr/bin/env python
"""Device-resident decode GB/s (k_dec_* kernels) of one library variant:  FCX_LIBFCX=... python tools/decode_dev.py zipf 64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from my_compress_b200 import api, datagen
family, mib = sys.argv[1], int(sys.argv[2])
n = mib << 20
nb = mib
data = datagen.gen(family, n, seed=20261018)
d_in = torch.from_numpy(data).cuda()
dc = api.DeviceCompressor(0, max_batch_blocks=min(1024, mib))
d_out, k = dc.compress(d_in)
blob = d_out[:k].cpu().numpy()
offs, end = api.record_offsets(blob, nb, 10)
d_file = torch.zeros(k + 64, dtype=torch.uint8, device="cuda"); d_file[:k] = d_out[:k]
d_dec = torch.empty(nb << 20, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2): lens = dc.decompress(d_file, offs, d_dec)
ok = int(lens.sum()) == n and bool(torch.equal(d_dec[:n], d_in))
ms = 0.0
reps = 5
for _ in range(reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dc.decompress(d_file, offs, d_dec); b.record(); torch.cuda.synchronize()
    ms += a.elapsed_time(b)
ms /= reps
print("%-22s %-8s %5d MiB  decode %.3f ms  %.2f GB/s  roundtrip %s" % (os.path.basename(os.environ.get("FCX_LIBFCX", "default")), family, mib, ms, n / ms / 1e6, "ok" if ok else "MISMATCH"))
