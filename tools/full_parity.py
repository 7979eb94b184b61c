#!/usr/bin/env python
"""Full-size parity of the BASELINE.json configs (run on the GPU box; test infrastructure).

    python tools/full_parity.py [config ...]        default: all that fit (configs 1,2,3,5 + 2 GiB of 4)

For each config: compress on the GPU, then
  * container structure: header fields, record sizes add up, every block parses
  * bit-exact parity against the oracle on a stride sample of blocks (host threads)
  * full round trip through the GPU decompressor (== input)
Prints one JSON line per config.
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _libs import oracle  # noqa: E402  (checker only)
from my_compress_b200 import api, datagen  # noqa: E402

CONFIGS = {
    "randm-1MiB": ("randm", 1 << 20, 1),
    "zipf-64MiB": ("zipf", 64 << 20, 64),
    "phrases-1GiB": ("phrases", 1 << 30, 48),
    "uniform-1GiB": ("uniform", 1 << 30, 48),
    "mixed-2GiB": ("mixed", 2 << 30, 48),          # config 4's corpus at a size one GPU takes comfortably
    "mixed-8GiB": ("mixed", 8 << 30, 96),          # config 4 itself: block-sharded over all visible GPUs
}


def check(name):
    family, nbytes, sample = CONFIGS[name]
    t0 = time.time()
    data = datagen.gen(family, nbytes, seed=20261018)
    t_gen = time.time() - t0
    t0 = time.time()
    n_gpus = max(1, api.device_count()) if name == "mixed-8GiB" else 1
    blob = api.compress(data, n_gpus=n_gpus)
    t_cmp = time.time() - t0
    nb = (nbytes + (1 << 20) - 1) >> 20
    assert blob[:4] == b"FCX7"
    assert int.from_bytes(blob[4:8], "little") == nbytes & 0xFFFFFFFF
    assert int.from_bytes(blob[8:10], "little") == nb & 0xFFFF
    offs, pos = [], 10
    for _ in range(nb):
        sz = int.from_bytes(blob[pos:pos + 4], "little")
        offs.append((pos + 4, sz))
        pos += 4 + sz
    assert pos == len(blob), "record sizes do not add up"
    idx = sorted(set(int(i * nb / sample) for i in range(min(sample, nb))))
    o = oracle()

    def one(b):
        want = o.compress_block(data[b << 20: min(nbytes, (b + 1) << 20)])
        s, z = offs[b]
        return blob[s:s + z] == want

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        ok = list(ex.map(one, idx))
    assert all(ok), "payload differs from the oracle in blocks %s" % [b for b, k in zip(idx, ok) if not k]
    t0 = time.time()
    back = api.decompress(blob)
    t_dec = time.time() - t0
    assert back == data.tobytes(), "round trip failed"
    return {"config": name, "bytes": nbytes, "blocks": nb, "container_bytes": len(blob), "ratio": len(blob) / nbytes,
            "oracle_parity_blocks": len(idx), "round_trip": "ok", "n_gpus": n_gpus, "gen_s": round(t_gen, 1),
            "compress_call_s": round(t_cmp, 2), "decompress_call_s": round(t_dec, 2)}


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for nm in names:
        print(json.dumps(check(nm)), flush=True)
