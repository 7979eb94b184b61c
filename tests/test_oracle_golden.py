"""Oracle port vs the committed golden vectors (tests/golden/golden.json, produced by
running the reference build -- see tests/golden/make_golden.py).  Runs anywhere."""
import hashlib
import json
import os

import numpy as np

from _libs import oracle
from my_compress_b200 import datagen

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_sunday_known_answers():
    # mc:3739-3759 in-source expectations: 15, 11, 15, -1, 4
    assert [c["pos"] for c in G["sunday"]] == [15, 11, 15, -1, 4]
    for c in G["sunday"]:
        assert oracle().sunday_search(c["main"].encode(), c["sub"].encode()) == c["pos"]


def test_golomb_and_pack_golden():
    for c in G["golomb"]:
        got = oracle().golomb_encode(np.array(c["in"], np.uint32))
        assert [int(w) for w in got] == c["words"]
    for c in G["combine11"]:
        assert oracle().combine_bits(np.array(c["in"], np.uint32)).hex() == c["hex"]


def test_huffman_golden():
    w = np.zeros(256, np.uint32)
    w[1:10] = G["huffman_tree"]["weights_1_to_9"]
    R, child = oracle().huffman_tree(w)
    assert R == G["huffman_tree"]["R"]
    for idx, node in G["huffman_tree"]["internal"].items():
        j = int(idx) - (512 - R)
        assert [int(child[2 * j]), int(child[2 * j + 1])] == node[2:4]
    for c in G["huffman_section"]:
        s = np.frombuffer(bytes.fromhex(c["in_hex"]), np.uint8)
        assert oracle().huffman_section(s).hex() == c["out_hex"]


def test_tokens_and_files_golden():
    t = G["lz77_tokens"]
    p, l, c = oracle().lz77_parse(np.frombuffer(t["in"].encode(), np.uint8))
    assert p.tolist() == t["p"] and l.tolist() == t["l"] and bytes(c).decode() == t["c"]
    f = G["file_small"]
    assert oracle().compress_file(np.frombuffer(f["in"].encode(), np.uint8)).hex() == f["fcx7_hex"]
    assert oracle().compress_file(np.full(5000, 65, np.uint8)).hex() == G["file_A5000"]["fcx7_hex"]
    assert oracle().compress_file(np.zeros(0, np.uint8)).hex() == G["file_empty"]["fcx7_hex"]


def test_block_payloads_golden():
    for b in G["blocks"]:
        x = datagen.gen_block(b["family"], b["n"], b["seed"], 0)
        assert hashlib.sha256(x.tobytes()).hexdigest() == b["in_sha256"], "datagen drifted: regenerate golden"
        out = oracle().compress_block(x)
        assert len(out) == b["payload_bytes"], b
        assert hashlib.sha256(out).hexdigest() == b["payload_sha256"], b
        dec = oracle().decompress_block(out)
        assert (dec == x.tobytes()) == b["ref_decodes_to_input"], b   # incl. the single-symbol defect (App. B1)


def test_two_block_container_golden():
    f = G["file_two_blocks"]
    data = np.concatenate([datagen.phrases_block(1 << 20, 3, 0), datagen.zipf_block(70000, 3, 1)])
    assert hashlib.sha256(data.tobytes()).hexdigest() == f["in_sha256"]
    blob = oracle().compress_file(data)
    assert len(blob) == f["bytes"] and hashlib.sha256(blob).hexdigest() == f["sha256"]
    assert oracle().decompress_file(blob, len(data)) == data.tobytes()


def test_lz78_golden():
    """LZ78 / FCX8 (scope row f-3): golden_lz78.json comes from the reference build (make_golden_lz78.py)."""
    L = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_lz78.json")))
    o = oracle()
    for t in L["tokens"]:
        idx, c = o.lz78_tokenize(np.frombuffer(bytes.fromhex(t["in_hex"]), np.uint8))
        assert idx.tolist() == t["idx"] and c.tobytes().hex() == t["c_hex"], t["in_hex"]
    for t in L["payload_hex"]:
        p = o.compress_block_lz78(np.frombuffer(bytes.fromhex(t["in_hex"]), np.uint8))
        assert p.hex() == t["fcx8_block_hex"], t["in_hex"]
        assert o.decompress_block_lz78(p).hex() == t["ref_decodes_hex"], t["in_hex"]
    for b in L["blocks"]:
        x = datagen.gen_block(b["family"], b["n"], b["seed"], 0)
        assert hashlib.sha256(x.tobytes()).hexdigest() == b["in_sha256"], "datagen drifted: regenerate golden"
        idx, _ = o.lz78_tokenize(x)
        assert len(idx) == b["tokens"] and hashlib.sha256(idx.tobytes()).hexdigest() == b["idx_sha256"], b
        p = o.compress_block_lz78(x)
        assert len(p) == b["payload_bytes"] and hashlib.sha256(p).hexdigest() == b["payload_sha256"], b
        d = o.decompress_block_lz78(p)
        assert len(d) == b["ref_decoded_bytes"] and (d == x.tobytes()) == b["ref_decodes_to_input"], b
