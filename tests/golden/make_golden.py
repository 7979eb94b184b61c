#!/usr/bin/env python
"""Generates tests/golden/golden.json by RUNNING THE REFERENCE (oracle/_ref/libmcref.so,
compiled from /root/reference by oracle/build_ref.sh).  Run here, commit the output:
/root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py
"""
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


def main():
    r = reference()
    g = {"generator": "tests/golden/make_golden.py via oracle/_ref/libmcref.so (unmodified reference functions)"}

    main_s = b"BBC ABCDAB ABCDABCDABDE"
    g["sunday"] = [
        {"main": main_s.decode(), "sub": s.decode(), "pos": r.sunday_search(main_s, s)}
        for s in (b"ABCDABD", b"ABCDABCDABDE", b"ABCDABDE", b"ABCDABDEF", b"ABCDAB")
    ]
    v = (np.arange(10) * 17).astype(np.uint32)
    g["golomb"] = [{"in": v.tolist(), "words": [int(w) for w in r.golomb_encode(v)]}]
    v = np.full(5, 257, np.uint32)
    g["golomb"].append({"in": v.tolist(), "words": [int(w) for w in r.golomb_encode(v)]})
    g["combine11"] = [
        {"in": [3, 3, 12, 5, 19], "hex": r.combine_bits(np.array([3, 3, 12, 5, 19], np.uint32)).hex()},
        {"in": [2047, 1, 1024, 2047], "hex": r.combine_bits(np.array([2047, 1, 1024, 2047], np.uint32)).hex()},
        {"in": [], "hex": r.combine_bits(np.zeros(0, np.uint32)).hex()},
    ]
    w = np.zeros(256, np.uint32)
    w[1:10] = [5, 29, 7, 0, 8, 14, 23, 3, 11]
    R, nodes = r.huffman_tree(w)
    g["huffman_tree"] = {"weights_1_to_9": w[1:10].tolist(), "R": int(R),
                         "internal": {str(i): [int(t) for t in nodes[i]] for i in range(512 - R, 511)}}
    g["huffman_section"] = []
    for s in (b"aaba", b"A" * 9, b"\x00", bytes(range(256)), b"abracadabra" * 7, b"\xff" * 7 + b"\x0f"):
        g["huffman_section"].append({"in_hex": s.hex(), "out_hex": r.huffman_section(np.frombuffer(s, np.uint8)).hex()})

    text = b"aacaacabcabaaacbaaacccaacabcad"
    p, l, c = r.lz77_parse(np.frombuffer(text, np.uint8))
    g["lz77_tokens"] = {"in": text.decode(), "p": p.tolist(), "l": l.tolist(), "c": bytes(c).decode()}
    g["file_small"] = {"in": text.decode(), "fcx7_hex": r.compress_file(np.frombuffer(text, np.uint8)).hex()}
    a5000 = r.compress_file(np.full(5000, 65, np.uint8))
    g["file_A5000"] = {"fcx7_hex": a5000.hex()}
    g["file_empty"] = {"fcx7_hex": r.compress_file(np.zeros(0, np.uint8)).hex()}

    # block payloads for seeded corpora: input is regenerated from datagen, output pinned by hash (+ size)
    g["blocks"] = []
    for fam, n, seed in (("randm", 65536, 1), ("uniform", 50000, 2), ("zipf", 65536, 3), ("phrases", 262144, 4),
                         ("zeros", 70000, 5), ("byterun", 9000, 6), ("periodic", 30000, 7),
                         ("zipf", 3, 8), ("zipf", 4, 9), ("phrases", 2047 + 300, 10)):
        x = datagen.gen_block(fam, n, seed, 0)
        out = r.compress_block(x)
        dec = r.decompress_block(out)
        g["blocks"].append({"family": fam, "n": n, "seed": seed,
                            "in_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
                            "payload_bytes": len(out), "payload_sha256": hashlib.sha256(out).hexdigest(),
                            "payload_head_hex": out[:48].hex(),
                            "ref_decodes_to_input": dec == x.tobytes()})
    # two-block container (phrases + ragged zipf tail)
    data = np.concatenate([datagen.phrases_block(1 << 20, 3, 0), datagen.zipf_block(70000, 3, 1)])
    blob = r.compress_file(data)
    g["file_two_blocks"] = {"in_sha256": hashlib.sha256(data.tobytes()).hexdigest(), "bytes": len(blob),
                            "sha256": hashlib.sha256(blob).hexdigest(), "head_hex": blob[:32].hex()}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
