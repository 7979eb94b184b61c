#!/usr/bin/env python
"""Generates tests/golden/golden_lz78.json from the REFERENCE build (oracle/_ref/libmcref.so, compiled from
/root/reference by oracle/build_ref.sh): LZ78 token lists, FCX8 block payloads and what the reference's own
decoder makes of them.  Run in the container that has /root/reference; the JSON is committed and
tests/test_oracle_golden.py checks the oracle against it anywhere."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _libs import reference  # noqa: E402
from my_compress_b200 import datagen  # noqa: E402

r = reference()
G = {"tokens": [], "payload_hex": [], "blocks": []}
for text in ("abababababa", "aaaaaaaaaaaaaaaa", "the cat sat on the mat, the cat sat", "abcabcabd\0", "x"):
    b = np.frombuffer(text.encode("latin1"), np.uint8)
    idx, c = r.lz78_tokenize(b)
    G["tokens"].append({"in_hex": b.tobytes().hex(), "idx": idx.tolist(), "c_hex": c.tobytes().hex()})
    p = r.compress_block_lz78(b)
    d = r.decompress_block_lz78(p)
    G["payload_hex"].append({"in_hex": b.tobytes().hex(), "fcx8_block_hex": p.hex(), "ref_decodes_hex": d.hex()})
for fam, n, seed in (("zipf", 300000, 2), ("phrases", 1 << 20, 2), ("uniform", 200000, 2), ("zipf", 1 << 20, 3), ("byterun", 70000, 4)):
    x = datagen.gen_block(fam, n, seed, 0)
    idx, c = r.lz78_tokenize(x)
    p = r.compress_block_lz78(x)
    d = r.decompress_block_lz78(p)
    G["blocks"].append({"family": fam, "n": n, "seed": seed, "in_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
                        "tokens": len(idx), "idx_sha256": hashlib.sha256(idx.tobytes()).hexdigest(),
                        "payload_bytes": len(p), "payload_sha256": hashlib.sha256(p).hexdigest(),
                        "ref_decodes_to_input": d == x.tobytes(), "ref_decoded_bytes": len(d)})
json.dump(G, open(os.path.join(HERE, "golden_lz78.json"), "w"), indent=1)
print("wrote golden_lz78.json:", len(G["tokens"]), "token lists,", len(G["blocks"]), "blocks")
