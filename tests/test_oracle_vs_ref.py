"""Pins the oracle port (oracle/fcx_oracle.c) against the reference's own functions
(oracle/_ref/libmcref.so, compiled from /root/reference by oracle/build_ref.sh).
Bit-exact, integer work: no tolerance."""
import numpy as np
import pytest

from _cases import small_buffers, stream_buffers
from _libs import oracle, reference

pytestmark = pytest.mark.ref


def test_sunday_known_answers_mc3739():
    # the only expected values the reference source itself carries (mc:3737-3759)
    cases = [
        (b"BBC ABCDAB ABCDABCDABDE", b"ABCDABD", 15),
        (b"BBC ABCDAB ABCDABCDABDE", b"ABCDABCDABDE", 11),
        (b"BBC ABCDAB ABCDABCDABDE", b"ABCDABDE", 15),
        (b"BBC ABCDAB ABCDABCDABDE", b"ABCDABDEF", -1),
        (b"BBC ABCDAB ABCDABCDABDE", b"ABCDAB", 4),
    ]
    for main, sub, _ in cases:
        assert oracle().sunday_search(main, sub) == reference().sunday_search(main, sub)


def test_match_rule_every_position():
    o, r = oracle(), reference()
    npos = 0
    for name, buf in small_buffers():
        po, lo = o.match_all(buf)
        pr, lr = r.match_all(buf)
        assert np.array_equal(lo, lr), name
        assert np.array_equal(po, pr), name
        npos += len(buf)
    assert npos > 100000


def test_match_bruteforce_agrees_with_chains():
    o = oracle()
    for name, buf in small_buffers(seed=5, count=16, max_len=3000):
        p, l = o.match_all(buf)
        for i in range(0, len(buf), 7):
            assert (p[i], l[i]) == o.longest_match_bf(buf, i), (name, i)


def test_parse_tokens():
    o, r = oracle(), reference()
    for name, buf in small_buffers(seed=77):
        a = o.lz77_parse(buf)
        b = r.lz77_parse(buf)
        for u, v in zip(a, b):
            assert np.array_equal(u, v), name


def test_huffman_tree_nodes():
    o, r = oracle(), reference()
    rng = np.random.default_rng(3)
    weight_sets = []  # all-zero weights never reach create_huffman_tree (mc:989 returns first)
    w = np.zeros(256, np.uint32); w[1:10] = [5, 29, 7, 0, 8, 14, 23, 3, 11]
    weight_sets.append(w)
    weight_sets.append(np.ones(256, np.uint32))
    weight_sets.append(np.full(256, 4096, np.uint32))
    for _ in range(200):
        k = int(rng.integers(1, 257))
        w = np.zeros(256, np.uint32)
        idx = rng.permutation(256)[:k]
        mode = rng.integers(0, 3)
        if mode == 0:
            w[idx] = rng.integers(1, 5, size=k)
        elif mode == 1:
            w[idx] = rng.integers(1, 100000, size=k)
        else:
            w[idx] = 1 << rng.integers(0, 12, size=k)
        weight_sets.append(w)
    for w in weight_sets:
        R, child = o.huffman_tree(w)
        Rr, nodes = r.huffman_tree(w)
        assert R == Rr
        if R >= 2:
            base = 512 - R
            for j in range(R - 1):
                assert child[2 * j] == nodes[base + j][2]
                assert child[2 * j + 1] == nodes[base + j][3]


def test_huffman_section_bytes():
    o, r = oracle(), reference()
    for name, s in stream_buffers():
        assert o.huffman_section(s) == r.huffman_section(s), name


def test_golomb_and_pack():
    o, r = oracle(), reference()
    rng = np.random.default_rng(11)
    vecs = [np.arange(10, dtype=np.uint32) * 17, np.full(40, 257, np.uint32), np.array([3], np.uint32), np.zeros(0, np.uint32)]
    for _ in range(100):
        n = int(rng.integers(1, 400))
        vecs.append(rng.integers(3, 258, size=n).astype(np.uint32))
    for v in vecs:
        go = o.golomb_encode(v)
        gr = r.golomb_encode(v)
        assert np.array_equal(go, gr)
        if len(v):
            assert np.array_equal(r.golomb_decode(gr, len(v)), v)
    for _ in range(100):
        n = int(rng.integers(0, 300))
        v = rng.integers(1, 2048, size=n).astype(np.uint32)
        assert o.combine_bits(v) == r.combine_bits(v)


def test_block_payload_bytes():
    o, r = oracle(), reference()
    for name, buf in small_buffers(seed=2024):
        if len(buf) == 0:
            continue  # main() never hands an empty block to the codec (mc:4090)
        assert o.compress_block(buf) == r.compress_block(buf), name


def test_container_bytes_multi_block():
    o, r = oracle(), reference()
    from my_compress_b200 import datagen

    data = np.concatenate([
        datagen.phrases_block(1 << 20, 3, 0),
        datagen.zipf_block(70000, 3, 1),
    ])
    a = o.compress_file(data)
    b = r.compress_file(data)
    assert a == b
    assert a[:4] == b"FCX7" and int.from_bytes(a[8:10], "little") == 2
    assert o.compress_file(np.zeros(0, np.uint8)) == r.compress_file(np.zeros(0, np.uint8))


def test_decoder_matches_reference_including_defects():
    o, r = oracle(), reference()
    cases = [b"A" * 5000, b"\0" * 4000, b"x", bytes(range(64)), b"aacaacabcabaaacbaaacccaacabcad"]
    for name, buf in small_buffers(seed=31, count=24, max_len=4000):
        if len(buf):
            cases.append(buf.tobytes())
    for data in cases:
        payload = r.compress_block(np.frombuffer(data, np.uint8))
        want = r.decompress_block(payload)
        got = o.decompress_block(payload)
        assert got == want, data[:16]


def test_lz78_tokens_payloads_and_decoder():
    """LZ78 / FCX8 (scope row f-3, "next"): the oracle's trie tokenizer, block payload and block decoder
    against the reference's hash-map tokenizer (mc:1832-1899), my_compress_file_lz78 (mc:3127-3476) and
    my_decompress_file_lz78 (mc:3478-3710), incl. the dropped trailing 0x00 (App. B2)."""
    from my_compress_b200 import datagen

    o, r = oracle(), reference()
    bufs = [(n, b) for n, b in small_buffers(seed=7) if len(b) > 0]
    bufs.append(("zipf300k", datagen.zipf_block(300000, 2, 0)))          # > 256 used indices: group Huffman tree
    bufs.append(("phrases200k", datagen.phrases_block(200000, 2, 0)))
    bufs.append(("uniform200k", datagen.uniform_block(200000, 2, 0)))
    bufs.append(("zeros70k", np.zeros(70000, np.uint8)))                 # ends inside a known phrase, decodes one 0x00 short
    bufs.append(("endzero", np.frombuffer(b"abcabcabd\0", np.uint8)))
    ntok = 0
    for name, buf in bufs:
        io, co = o.lz78_tokenize(buf)
        ir, cr = r.lz78_tokenize(buf)
        assert np.array_equal(io, ir) and np.array_equal(co, cr), name
        ntok += len(io)
        po, pr = o.compress_block_lz78(buf), r.compress_block_lz78(buf)
        assert po == pr, name
        assert o.decompress_block_lz78(pr) == r.decompress_block_lz78(pr), name
    assert ntok > 150000


def test_lz78_container_is_what_the_reference_cli_writes(tmp_path):
    """FCX8 file level: the oracle's container equals the file the reference CLI writes with `-c lz78`
    (mc:4073-4136), and both decoders give the same bytes back (one trailing 0x00 per block may be lost, App. B2)."""
    import os
    import subprocess

    from _libs import REF_CLI
    from my_compress_b200 import datagen

    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/my_compress_ref not built")
    data = np.concatenate([datagen.zipf_block(1 << 20, 9, 0), datagen.phrases_block(1 << 20, 9, 1), datagen.uniform_block(12345, 9, 2)])
    src, cmp_, back = tmp_path / "in.bin", tmp_path / "out.fcx8", tmp_path / "back.bin"
    src.write_bytes(data.tobytes())
    subprocess.check_call([REF_CLI, "-i", str(src), "-o", str(cmp_), "-c", "lz78"], stdout=subprocess.DEVNULL)
    blob = cmp_.read_bytes()
    assert blob[:4] == b"FCX8"
    assert oracle().compress_file_lz78(data) == blob
    subprocess.check_call([REF_CLI, "-i", str(cmp_), "-o", str(back)], stdout=subprocess.DEVNULL)
    assert oracle().decompress_file(blob, len(data) + 64) == back.read_bytes() == data.tobytes()
