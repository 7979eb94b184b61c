"""GPU parity: the CUDA path (through the C ABI in libfcx.so) against the oracle port,
stage by stage and end to end.  Integer/byte work: bit-exact, no tolerance."""
import hashlib
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from _cases import small_buffers, stream_buffers
from _libs import oracle
from my_compress_b200 import api, datagen

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
GL = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_lz78.json")))


@pytest.fixture(scope="module")
def dc():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    d = api.DeviceCompressor(0, max_batch_blocks=8)
    yield d
    d.close()


def _dev(buf):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8)).cuda()
    pad = torch.zeros(((t.numel() + 15) // 16) * 16 + 16, dtype=torch.uint8, device="cuda")
    pad[: t.numel()] = t
    return pad[: t.numel()]


def _first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n


def test_match_every_position(dc):
    o = oracle()
    bufs = [(n, b) for n, b in small_buffers() if len(b) > 0]
    bufs.append(("zipf64k", datagen.zipf_block(65536, 1, 0)))
    bufs.append(("phrases100k", datagen.phrases_block(100000, 1, 0)))
    bufs.append(("zeros20k", np.zeros(20000, np.uint8)))
    bufs.append(("uniform30k", datagen.uniform_block(30000, 1, 0)))
    for name, buf in bufs:
        po, lo = o.match_all(buf)
        lg, pg = dc.probe_match(_dev(buf))
        bad = np.nonzero((lg != lo) | (pg != po))[0]
        assert len(bad) == 0, "%s: first mismatch at %d: gpu (p=%d,l=%d) oracle (p=%d,l=%d)" % (
            name, bad[0], pg[bad[0]], lg[bad[0]], po[bad[0]], lo[bad[0]])


def test_match_auto_mode_is_exact_wherever_a_parse_can_go(dc):
    """Product-path run of the match kernel: tiles of long matches are evaluated lazily, so a
    position may read 0 -- but never on the parse chain, and an evaluated position is exact."""
    o = oracle()
    bufs = [("phrases300k", datagen.phrases_block(300000, 3, 0)), ("zeros50k", np.zeros(50000, np.uint8)),
            ("byterun9000", datagen.byte_run_block(9000, 6, 0)), ("periodic70k", datagen.periodic_block(70000, 4, 2)),
            ("zipf64k", datagen.zipf_block(65536, 1, 0)),
            ("mix", np.concatenate([datagen.phrases_block(20000, 5, 0), datagen.uniform_block(9000, 5, 1),
                                    datagen.phrases_block(30000, 5, 2), datagen.zipf_block(9000, 5, 3)]))]
    lazy_seen = False
    for name, buf in bufs:
        po, lo = o.match_all(buf)
        lg, pg = dc.probe_match(_dev(buf), auto=True)
        ev = lg != 0
        assert np.array_equal(lg[ev], lo[ev]) and np.array_equal(pg[ev], po[ev]), name
        i, chain = 0, []
        while i < len(buf):
            chain.append(i)
            i += int(lo[i]) + 1
        chain = np.array(chain)
        assert np.array_equal(lg[chain], lo[chain]) and np.array_equal(pg[chain], po[chain]), name
        lazy_seen |= bool(np.any((lg == 0) & (lo != 0)))
    assert lazy_seen, "no tile took the lazy path: the test inputs no longer exercise it"


def _oracle_streams(buf):
    o = oracle()
    p, l, c = o.lz77_parse(buf)
    T = len(c)
    bm = np.zeros((T + 7) // 8, np.uint8)
    lit = np.nonzero(l == 0)[0]
    np.bitwise_or.at(bm, lit >> 3, (1 << (lit & 7)).astype(np.uint8))
    m = l != 0
    gw = o.golomb_encode(l[m])
    return {"T": T, "Pc": int(m.sum()), "G": len(gw), "bitmap": bm.tobytes(), "c": c.tobytes(),
            "p": o.combine_bits(p[m]), "golomb": gw.tobytes()}


def test_parse_and_stream_split(dc):
    bufs = [(n, b) for n, b in small_buffers(seed=4321) if len(b) > 0]
    bufs.append(("zipf300k", datagen.zipf_block(300000, 2, 0)))
    bufs.append(("phrases1M", datagen.phrases_block(1 << 20, 2, 0)))
    bufs.append(("uniform70k", datagen.uniform_block(70000, 2, 0)))
    bufs.append(("byterun9000", datagen.byte_run_block(9000, 6, 0)))
    for name, buf in bufs:
        want = _oracle_streams(buf)
        got = dc.probe_streams(_dev(buf))
        for k in ("T", "Pc", "G"):
            assert got[k] == want[k], (name, k, got[k], want[k])
        for k in ("bitmap", "c", "p", "golomb"):
            assert got[k] == want[k], "%s: stream %s differs at byte %d" % (name, k, _first_diff(got[k], want[k]))


def test_huffman_section_bytes(dc):
    o = oracle()
    bufs = list(stream_buffers())
    bufs.append(("uniform1M", datagen.uniform_block(1 << 20, 7, 0)))
    bufs.append(("zipf500001", datagen.zipf_block(500001, 7, 0)))
    bufs.append(("bits8192", np.random.default_rng(5).integers(0, 2, size=8192, dtype=np.uint8)))
    bufs.append(("bits8193", np.random.default_rng(5).integers(0, 2, size=8193, dtype=np.uint8)))
    bufs.append(("skew24577", np.minimum(np.random.default_rng(6).geometric(0.5, size=24577) - 1, 255).astype(np.uint8)))
    for name, s in bufs:
        want = o.huffman_section(s)
        got = dc.probe_huffman_section(_dev(s))
        assert got == want, "%s: section differs at byte %d of %d/%d" % (name, _first_diff(got, want), len(got), len(want))


def test_huffman_encode_groups_and_chunk_edges(dc):
    """k_encode_w picks 4 / 3 / 2 / 1 symbols per 32-bit insert by the stream's longest code (<= 8 / 10 /
    16 / 24 bits; longer goes to k_encode_slow): Fibonacci weights over m symbols give a longest code
    of m-1 bits.  Lengths sit around the 32-symbol run and 8192-symbol chunk boundaries, where the
    ragged tail, the bulk-copy size and the descriptor batch change behaviour."""
    o = oracle()
    rng = np.random.default_rng(77)
    for m in (2, 9, 10, 11, 12, 17, 18, 25, 26, 28):
        fib = [1, 1]
        while len(fib) < m:
            fib.append(fib[-1] + fib[-2])
        base = rng.permutation(np.repeat(np.arange(m, dtype=np.uint8) * 7 + 3, fib))
        while len(base) < 8192 * 3 + 40:                      # short multisets: repeat (the weights keep their ratios)
            base = np.concatenate([base, rng.permutation(base)])
        F = min(len(base), 1 << 20)
        sizes = {F, F - 1, 31, 32, 33, 63, 8191, 8192, 8193, 8192 + 31, 8192 * 2 + 32, 8192 * 3 + 1}
        for n in sorted(x for x in sizes if 0 < x <= F):
            s = np.ascontiguousarray(base[:n])
            want = o.huffman_section(s)
            got = dc.probe_huffman_section(_dev(s))
            assert got == want, "m=%d n=%d: section differs at byte %d of %d/%d" % (m, n, _first_diff(got, want), len(got), len(want))


def test_huffman_encode_fixed_8bit_streams(dc):
    """A flat histogram over all 256 bytes gives every symbol an 8-bit code; k_encode_w then translates
    bytes instead of packing bits (EW_FIXED8).  Sizes sit around its units: 4 symbols per word, 1024 per
    warp, 8192 per chunk, and the ragged 1..3 bytes of a stream end; the section's byte offset inside the
    probe container varies with the header size, so all four word alignments occur across the cases."""
    o = oracle()
    rng = np.random.default_rng(123)
    for n in (1024, 1025, 1026, 1027, 2047, 2048, 4099, 8191, 8192, 8193, 8192 + 1023, 8192 + 1025, 3 * 8192 + 2, 65536 + 7):
        reps = (n + 255) // 256
        s = np.concatenate([rng.permutation(256) for _ in range(reps)]).astype(np.uint8)[:n]
        want = o.huffman_section(s)
        got = dc.probe_huffman_section(_dev(s))
        assert got == want, "n=%d: section differs at byte %d of %d/%d" % (n, _first_diff(got, want), len(got), len(want))
    # as the c stream of a block, where the section starts at whatever byte the bitmap section ends on
    for n in (70001, 70002, 70003, 70004):
        x = datagen.uniform_block(n, 31, 0)
        assert api.my_compress_file_lz77(x) == o.compress_block(x), n


def test_encode_direct_path_for_dense_subchunks():
    """The small-stage build (my_compress_b200/build.py build_test_variants: 96 stage words per warp) sends
    every sub-chunk denser than ~3 bits per symbol down k_encode_w's direct path; the library is chosen at
    import time, so the check runs in its own process."""
    import subprocess
    import sys

    from my_compress_b200 import build as b

    if not os.path.exists(b.LIB_SMALLSTAGE):
        pytest.skip("libfcx_smallw.so not built (run __graft_entry__.build())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FCX_LIBFCX=b.LIB_SMALLSTAGE)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_variant.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_block_payloads_vs_oracle_and_golden():
    o = oracle()
    for name, buf in small_buffers(seed=2024):
        if len(buf) == 0:
            continue
        want = o.compress_block(buf)
        got = api.my_compress_file_lz77(buf)
        assert got == want, "%s: payload differs at byte %d" % (name, _first_diff(got, want))
    for b in G["blocks"]:
        x = datagen.gen_block(b["family"], b["n"], b["seed"], 0)
        got = api.my_compress_file_lz77(x)
        assert len(got) == b["payload_bytes"] and hashlib.sha256(got).hexdigest() == b["payload_sha256"], b


def test_golden_files():
    f = G["file_small"]
    assert api.compress(f["in"].encode()).hex() == f["fcx7_hex"]
    assert api.compress(np.full(5000, 65, np.uint8)).hex() == G["file_A5000"]["fcx7_hex"]
    assert api.compress(b"").hex() == G["file_empty"]["fcx7_hex"]
    t = G["file_two_blocks"]
    data = np.concatenate([datagen.phrases_block(1 << 20, 3, 0), datagen.zipf_block(70000, 3, 1)])
    blob = api.compress(data)
    assert len(blob) == t["bytes"] and hashlib.sha256(blob).hexdigest() == t["sha256"]


def _oracle_container(data):
    """Oracle container with the per-block work fanned out over host threads (ctypes drops the GIL)."""
    o = oracle()
    n = len(data)
    nb = (n + (1 << 20) - 1) >> 20
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        payloads = list(ex.map(lambda b: o.compress_block(data[b << 20 : min(n, (b + 1) << 20)]), range(nb)))
    head = b"FCX7" + (n & 0xFFFFFFFF).to_bytes(4, "little") + (nb & 0xFFFF).to_bytes(2, "little")
    return head + b"".join(len(p).to_bytes(4, "little") + p for p in payloads)


@pytest.mark.parametrize("family,nbytes", [
    ("randm", 1 << 20),                  # config 1
    ("zipf", (8 << 20) + 12345),         # config 2 family, ragged tail
    ("phrases", 16 << 20),               # config 3 family
    ("mixed", 12 << 20),                 # config 4 family
    ("uniform", (4 << 20) + 1),          # config 5 family, 1-byte last block
    ("byterun", (1 << 20) + 5000),       # worst case for the match finder
    ("periodic", 3 << 20),
    ("mixed", (40 << 20) + 5),           # > 16 blocks: the host path runs as a pipeline of sub-batches
])
def test_container_vs_oracle(family, nbytes):
    data = datagen.gen(family, nbytes, seed=11)
    want = _oracle_container(data)
    got = api.compress(data)
    assert len(got) == len(want), (len(got), len(want))
    assert got == want, "container differs at byte %d" % _first_diff(got, want)


def test_decompress_matches_reference_behaviour():
    o = oracle()
    cases = [b"A" * 5000, b"\0" * 4000, b"x", bytes(range(64)), b"aacaacabcabaaacbaaacccaacabcad"]
    for name, buf in small_buffers(seed=31, count=24, max_len=4000):
        if len(buf):
            cases.append(buf.tobytes())
    for data in cases:
        payload = o.compress_block(np.frombuffer(data, np.uint8))
        want = o.decompress_block(payload)
        got = api.my_decompress_file_lz77(payload)
        assert got == want, data[:16]


def test_round_trip_container():
    for family, nbytes in (("zipf", (2 << 20) + 77), ("phrases", 3 << 20), ("uniform", 1 << 20)):
        data = datagen.gen(family, nbytes, seed=5)
        blob = api.compress(data)
        assert api.decompress(blob) == data.tobytes()


def test_round_trip_hard_streams_for_the_decoder():
    """Long Huffman codes (tree walk beyond the 11-bit table), Golomb runs of 64 ones (l = 257), periodic
    overlapping copies, a ragged last block: the speculative decoder must land on the sequential result,
    i.e. on what the reference's decoder makes of the same container."""
    rng = np.random.default_rng(4242)
    geo = np.minimum(rng.geometric(0.45, size=(1 << 20) + 4097) - 1, 255).astype(np.uint8)     # codes up to ~20 bits
    cases = [("geometric", geo),
             ("byterun", datagen.gen("byterun", (2 << 20) + 33, seed=8)),
             ("periodic", datagen.gen("periodic", (1 << 20) + 1, seed=8)),
             ("randm", datagen.gen("randm", 1 << 20, seed=8)),
             ("two_symbols", (rng.integers(0, 2, size=300001, dtype=np.uint8) * 7 + 65).astype(np.uint8)),
             ("mixed", datagen.gen("mixed", (5 << 20) + 12345, seed=8))]
    o = oracle()
    for name, data in cases:
        blob = api.compress(data)
        want = o.decompress_file(blob, len(data) + 64)       # the reference's decoder (lossy on single-symbol streams, App. B1)
        got = api.decompress(blob)
        assert got == want, "%s: differs at byte %d of %d/%d" % (name, _first_diff(got, want), len(got), len(want))
        if name in ("geometric", "randm", "mixed"):          # no single-symbol stream in these: the round trip is exact
            assert got == data.tobytes(), name


def _section_end(p, at):
    tn = p[at]
    w_at = at + 1 + (2 * tn + 7) // 8 + 2 * tn
    return w_at + 4 + 4 * int.from_bytes(p[w_at:w_at + 4], "little")


def test_decoder_stops_where_the_reference_does_when_matches_run_out():
    """mc:2337-2340: a match flag with no distance left prints "Fatal Error" and ends the block early.
    Payloads with the stored match count lowered must decode to the same prefix as the reference."""
    o = oracle()
    for name, buf in small_buffers(seed=77, count=16, max_len=6000):
        if len(buf) < 64:
            continue
        p = bytearray(o.compress_block(buf))
        T = int.from_bytes(p[0:4], "little")
        if (T + 7) // 8 <= 1:
            continue
        at = _section_end(p, _section_end(p, 4))            # past the bitmap and c sections: the Pc field
        pc = int.from_bytes(p[at:at + 4], "little")
        for new_pc in {0, pc // 2, max(0, pc - 1)} - {pc}:
            q = bytearray(p)
            q[at:at + 4] = new_pc.to_bytes(4, "little")
            want = o.decompress_block(bytes(q))
            got = api.my_decompress_file_lz77(bytes(q))
            assert got == want, (name, pc, new_pc, len(got), len(want))


def test_multi_gpu_block_sharding_matches_single_gpu():
    """fcx_compress_lz77 with a multi-device mask shards contiguous block ranges over the GPUs;
    the container must not depend on the device count."""
    ndev = api.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    data = datagen.gen("mixed", (9 << 20) + 777, seed=21)
    one = api.compress(data, n_gpus=1)
    for g in sorted({2, min(ndev, 4), ndev}):
        assert api.compress(data, n_gpus=g) == one, "differs with %d GPUs" % g
    assert one == _oracle_container(data)


def test_cli_is_interchangeable_with_the_reference_cli(tmp_path):
    """./fcx_cli takes the reference's flags (mc:3992-4053) and writes the same files: compare with the
    reference CLI built by oracle/build_ref.sh, and cross-decompress in both directions."""
    import subprocess

    from _libs import REF_CLI
    from my_compress_b200 import build

    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/my_compress_ref not built")
    cli = build.build_cli()
    data = datagen.gen("mixed", (2 << 20) + 999, seed=4)
    src = tmp_path / "in.bin"
    src.write_bytes(data.tobytes())
    ours, ref = tmp_path / "ours.fcx7", tmp_path / "ref.fcx7"
    subprocess.check_call([cli, "-i", str(src), "-o", str(ours), "-c", "lz77"])
    subprocess.check_call([REF_CLI, "-i", str(src), "-o", str(ref), "-c", "lz77"], stdout=subprocess.DEVNULL)
    assert ours.read_bytes() == ref.read_bytes()
    back1, back2 = tmp_path / "back1.bin", tmp_path / "back2.bin"
    subprocess.check_call([REF_CLI, "-i", str(ours), "-o", str(back1)], stdout=subprocess.DEVNULL)   # reference decodes ours
    subprocess.check_call([cli, "-i", str(ref), "-o", str(back2)])                                     # we decode the reference's
    assert back1.read_bytes() == data.tobytes() and back2.read_bytes() == data.tobytes()
    # "-c <anything but lz77>" is lz78 (mc:4040-4043): FCX8 files, interchangeable both ways as well
    ours8, ref8, back8a, back8b = tmp_path / "ours.fcx8", tmp_path / "ref.fcx8", tmp_path / "back8a.bin", tmp_path / "back8b.bin"
    subprocess.check_call([cli, "-i", str(src), "-o", str(ours8), "-c", "lz78"])
    subprocess.check_call([REF_CLI, "-i", str(src), "-o", str(ref8), "-c", "lz78"], stdout=subprocess.DEVNULL)
    assert ours8.read_bytes() == ref8.read_bytes()
    subprocess.check_call([REF_CLI, "-i", str(ours8), "-o", str(back8a)], stdout=subprocess.DEVNULL)
    subprocess.check_call([cli, "-i", str(ref8), "-o", str(back8b)])
    assert back8a.read_bytes() == back8b.read_bytes()            # the reference's decoder output (it may drop a trailing 0x00)
    # the streaming loop proper: a build whose file pieces are 2 blocks reads / compresses / writes this 3-block
    # file in two overlapped pieces and rewrites the header at the end (mc:4128-4129); same bytes both ways
    small = build.CLI_SMALL
    ours2, back3 = tmp_path / "ours2.fcx7", tmp_path / "back3.bin"
    subprocess.check_call([small, "-i", str(src), "-o", str(ours2), "-c", "lz77"])
    assert ours2.read_bytes() == ref.read_bytes()
    subprocess.check_call([small, "-i", str(ref), "-o", str(back3)])
    assert back3.read_bytes() == data.tobytes()
    empty = tmp_path / "empty.bin"
    empty.write_bytes(b"")
    subprocess.check_call([small, "-i", str(empty), "-o", str(ours2), "-c", "lz77"])
    assert ours2.read_bytes() == bytes.fromhex(G["file_empty"]["fcx7_hex"])


def test_device_context_with_small_batches_and_capacity_error(dc):
    """A context whose batch (8 blocks here, see the fixture) is smaller than the input walks the blocks
    batch by batch with a device-side output cursor; an undersized output buffer is reported, not overrun."""
    import torch

    data = datagen.gen("mixed", (19 << 20) + 123, seed=8)            # 20 blocks -> 3 batches of <= 8
    d_in = _dev(data)
    want = _oracle_container(data)
    d_out = torch.zeros(api.compress_bound(len(data)), dtype=torch.uint8, device="cuda")
    _, n = dc.compress(d_in, d_out)
    assert bytes(d_out[:n].cpu().numpy()) == want
    small = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")    # far too small for 20 blocks
    with pytest.raises(api.FcxError) as e:
        dc.compress(d_in, small)
    assert e.value.code == -4                                          # FCX_ERR_CAPACITY
    assert bytes(small[-16:].cpu().numpy()) == bytes(16)               # nothing written past / near the end
    _, n2 = dc.compress(d_in, d_out)                                   # the context is still usable
    assert n2 == n


def test_records_api_and_large_pipelines_round_trip():
    """fcx_compress_records_lz77 is the block loop without the header: records of consecutive pieces behind a
    header are the file.  A 300-block input walks the multi-batch compress pipeline and the decoder's batch
    ring (batches of <= 128 blocks, three device slots) -- checked against the input, no CPU reference needed
    at this size (the per-block bytes are pinned by the other tests)."""
    data = datagen.gen("phrases", (300 << 20) + 4321, seed=3)
    cut = 77 << 20
    recs = api.compress_records(data[:cut]) + api.compress_records(data[cut:])
    nb = (len(data) + (1 << 20) - 1) >> 20
    head = b"FCX7" + (len(data) & 0xFFFFFFFF).to_bytes(4, "little") + (nb & 0xFFFF).to_bytes(2, "little")
    blob = api.compress(data)
    assert blob == head + recs
    assert api.decompress(blob) == data.tobytes()
    assert api.decompress_records(recs, nb) == data.tobytes()
    # pinned host buffers from the library's own allocator take the same path
    pin_in, pin_out = api.PinnedBuffer(len(data)), api.PinnedBuffer(api.compress_bound(len(data)))
    pin_in.array[:] = data
    n = api.compress_into(pin_in.array, pin_out.array)
    assert pin_out.array[:n].tobytes() == blob
    pin_in.close(); pin_out.close()


def test_device_resident_decode_and_truncated_containers(dc):
    import torch

    data = datagen.gen("mixed", (5 << 20) + 99, seed=13)
    blob = np.frombuffer(api.compress(data), np.uint8)
    offs, end = api.record_offsets(blob, 6, 10)
    assert end == len(blob)
    d_file = torch.zeros(len(blob) + 64, dtype=torch.uint8, device="cuda")
    d_file[: len(blob)] = torch.from_numpy(blob.copy()).cuda()
    d_out = torch.zeros(6 << 20, dtype=torch.uint8, device="cuda")
    lens = dc.decompress(d_file, offs, d_out)
    assert lens.tolist() == [1 << 20] * 5 + [99]
    assert bytes(d_out[: len(data)].cpu().numpy()) == data.tobytes()
    for cutoff in (5, 12, len(blob) // 2, len(blob) - 1):
        with pytest.raises(api.FcxError) as e:
            api.decompress(blob[:cutoff].tobytes() if cutoff >= 10 else blob[:cutoff].tobytes() + b"")
        assert e.value.code == -5                                      # FCX_ERR_FORMAT (mc:4139-4152: read errors end the run)


def test_no_kernel_writes_outside_its_workspace_arrays():
    """compute-sanitizer is not available on the GPU pool: every workspace array of a context lies between two
    guard bands; after the whole pipeline ran over every data family (ragged tail, lazy long-match tiles,
    literal-only tiles, multi-batch, decode) none of them may have changed."""
    import torch

    d = api.DeviceCompressor(0, max_batch_blocks=4)
    try:
        d.guards(False)
        for family, nbytes in (("mixed", (9 << 20) + 12345), ("byterun", (2 << 20) + 1), ("zeros", 5000), ("periodic", (1 << 20) + 7),
                               ("zipf", 1), ("uniform", (1 << 20) - 1)):
            data = datagen.gen(family, nbytes, seed=17)
            d_in = _dev(data)
            d_out = torch.empty(api.compress_bound(len(data)), dtype=torch.uint8, device="cuda")
            _, n = d.compress(d_in, d_out)
            blob = d_out[:n].cpu().numpy()
            nb = (nbytes + (1 << 20) - 1) >> 20
            offs, _ = api.record_offsets(blob, nb, 10)
            d_file = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
            d_file[:n] = d_out[:n]
            d.decompress(d_file, offs, torch.empty(nb << 20, dtype=torch.uint8, device="cuda"))
            assert d.guards(True) == 0, family
    finally:
        d.close()


def test_lz78_payloads_containers_and_decoder_vs_oracle():
    """FCX8 (scope row f-3): block payloads and containers byte for byte against the oracle's restatement of
    my_compress_file_lz78 (pinned against the reference in tests/test_oracle_vs_ref.py), the decoder against the
    reference decoder's behaviour -- including the unconditional drop of a trailing 0x00 (App. B2) and an input
    that ends inside a known phrase (mc:1858-1862)."""
    o = oracle()
    bufs = [(n, b) for n, b in small_buffers(seed=78, count=24, max_len=5000) if len(b) > 0]
    bufs.append(("abab", np.frombuffer(b"ABBCBCABABCAABCAAB", np.uint8)))
    bufs.append(("ends_in_phrase", np.frombuffer(b"abcabcabcabc" * 9 + b"abc", np.uint8)))
    bufs.append(("ends_with_zero", np.frombuffer(b"hello world\0", np.uint8)))
    bufs.append(("zipf300k", datagen.zipf_block(300000, 78, 0)))
    bufs.append(("uniform1M", datagen.uniform_block(1 << 20, 78, 1)))                # > 256 used ids: group tree + group codes
    bufs.append(("phrases1M", datagen.phrases_block(1 << 20, 78, 2)))
    bufs.append(("zeros200k", np.zeros(200000, np.uint8)))                           # phrase k has length k: deep parent chains
    for name, buf in bufs:
        want = o.compress_block_lz78(buf)
        got = api.my_compress_file_lz78(buf)
        assert got == want, "%s: payload differs at byte %d of %d/%d" % (name, _first_diff(got, want), len(got), len(want))
        dec_want = o.decompress_block_lz78(want)
        dec_got = api.my_decompress_file_lz78(want)
        assert dec_got == dec_want, "%s: decoder differs at byte %d of %d/%d" % (name, _first_diff(dec_got, dec_want), len(dec_got), len(dec_want))
    data = np.concatenate([datagen.zipf_block(1 << 20, 5, 0), datagen.uniform_block(1 << 20, 5, 1), datagen.phrases_block(70001, 5, 2)])
    blob = api.compress_lz78(data)
    assert blob == o.compress_file_lz78(data)
    assert api.decompress(blob) == o.decompress_file(blob, len(data) + 64)
    for g in GL["payload_hex"]:                                                       # committed vectors made by the reference build
        x = np.frombuffer(bytes.fromhex(g["in_hex"]), np.uint8)
        assert api.my_compress_file_lz78(x).hex() == g["fcx8_block_hex"]
        assert api.my_decompress_file_lz78(bytes.fromhex(g["fcx8_block_hex"])).hex() == g["ref_decodes_hex"]
