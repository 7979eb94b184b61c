"""FCX9, the opt-in format of scope row f-4 (include/fcx.h): 16-byte header with u64 bytes / u32 blocks, and a Huffman
section without tree nodes names its symbol.  The reference has no such format, so the checker is the FCX7
restatement with those two deltas; these tests tie it back to FCX7 byte by byte and check what the format exists
for -- decode(encode(x)) == x on the inputs where the reference's own round trip fails.
CPU part: the oracle.  GPU part (marked gpu): the CUDA path through the C ABI against it."""
import numpy as np
import pytest

from _cases import small_buffers
from _libs import oracle
from my_compress_b200 import datagen

B1_CASES = [b"A" * 5000, b"x", bytes(range(64)), b"\0" * 4000, b"ab" * 3000, b"q" * 7]   # the first three fail in FCX7


def _sections(payload, x9):
    """Walks a block payload (mc:2115-2253); returns [(offset of the section's first byte, Tn)] of its Huffman sections."""
    rd = lambda at: int.from_bytes(payload[at:at + 4], "little")
    out = []
    T = rd(0)
    at = 4

    def section(at):
        Tn = payload[at]
        out.append((at, Tn))
        named = 1 if (x9 and Tn == 0) else 0
        w_at = at + 1 + named + (2 * Tn + 7) // 8 + 2 * Tn
        return w_at + 4 + 4 * rd(w_at)

    map_b = (T + 7) // 8
    at = section(at) if map_b > 1 else at + map_b
    at = section(at)
    at += 4                                  # Pc
    at = section(at)
    G = rd(at)
    at += 4
    if G:
        at = section(at)
    assert at == len(payload)
    return out


def _strip_names(p9):
    """FCX9 payload -> the FCX7 payload of the same block: drop the symbol byte behind every zero node count."""
    drop = sorted(at + 1 for at, Tn in _sections(p9, True) if Tn == 0)
    b = bytearray(p9)
    for at in reversed(drop):
        del b[at]
    return bytes(b), len(drop)


def test_oracle_x9_payload_is_fcx7_plus_the_named_symbols():
    o = oracle()
    named_total = 0
    for name, buf in small_buffers(seed=77, count=40, max_len=5000):
        if len(buf) == 0:
            continue
        p7, p9 = o.compress_block(buf), o.compress_block_x9(buf)
        back, named = _strip_names(p9)
        assert back == p7, name
        assert len(p9) == len(p7) + named
        named_total += named
        assert o.decompress_block_x9(p9) == buf.tobytes(), name          # exact round trip, whatever the input
    assert named_total > 0                                               # the cases do reach the escape


def test_oracle_x9_fixes_the_round_trips_the_reference_format_loses():
    o = oracle()
    lossy = 0
    for data in B1_CASES:
        x = np.frombuffer(data, np.uint8)
        assert o.decompress_block_x9(o.compress_block_x9(x)) == data
        if o.decompress_block(o.compress_block(x)) != data:
            lossy += 1
    assert lossy >= 3                                                    # App. B1 is real in FCX7


def test_oracle_x9_container_header_and_round_trip():
    o = oracle()
    data = np.concatenate([datagen.zipf_block(1 << 20, 5, 0), np.full(70000, 65, np.uint8)])
    blob = o.compress_file_x9(data)
    assert blob[:4] == b"FCX9"
    assert int.from_bytes(blob[4:12], "little") == len(data) and int.from_bytes(blob[12:16], "little") == 2
    assert o.decompress_file(blob, len(data) + 16) == data.tobytes()
    assert o.compress_file_x9(b"") == b"FCX9" + bytes(12)
    # block 0 has no single-symbol stream: its record is the FCX7 record
    b7 = o.compress_file(data)
    n0 = int.from_bytes(b7[10:14], "little")
    assert blob[16:16 + 4 + n0] == b7[10:10 + 4 + n0]


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_x9_payloads_and_decoder_vs_oracle():
    from my_compress_b200 import api

    o = oracle()
    for name, buf in small_buffers(seed=78, count=32, max_len=5000):
        if len(buf) == 0:
            continue
        want = o.compress_block_x9(buf)
        got = api.compress_block_x9(buf)
        assert got == want, name
        assert api.decompress_block_x9(got) == buf.tobytes(), name
    for data in B1_CASES:
        x = np.frombuffer(data, np.uint8)
        p = api.compress_block_x9(x)
        assert p == o.compress_block_x9(x)
        assert api.decompress_block_x9(p) == data


@pytest.mark.gpu
def test_gpu_x9_containers_round_trip_exactly():
    from my_compress_b200 import api

    o = oracle()
    assert api.compress_x9(b"") == b"FCX9" + bytes(12)
    assert api.decompress(api.compress_x9(b"")) == b""
    cases = [np.frombuffer(d, np.uint8) for d in B1_CASES]
    cases.append(np.concatenate([datagen.zipf_block(1 << 20, 5, 0), np.full(70000, 65, np.uint8)]))
    cases.append(datagen.gen("mixed", (20 << 20) + 3, seed=9))           # > 16 blocks: the pipelined host path
    cases.append(datagen.gen("byterun", (2 << 20) + 100, seed=9))        # every section single-symbol: lossy in FCX7
    for data in cases:
        blob = api.compress_x9(data)
        if len(data) <= (3 << 20):
            assert blob == o.compress_file_x9(data)
        assert blob[:4] == b"FCX9" and int.from_bytes(blob[4:12], "little") == len(data)
        assert api.decompress(blob) == data.tobytes()
    # records without a header, both directions
    data = datagen.gen("byterun", (1 << 20) + 9, seed=2)
    rec = api.compress_records_x9(data)
    assert rec == o.compress_file_x9(data)[16:]
    assert api.decompress_records_x9(rec, 2) == data.tobytes()


@pytest.mark.gpu
def test_gpu_x9_device_resident_path():
    import torch
    from my_compress_b200 import api

    data = np.concatenate([datagen.gen("phrases+run", 3 << 20, seed=4), np.full(5000, 7, np.uint8)])
    d_in = torch.from_numpy(data).cuda()
    dc = api.DeviceCompressor(0, max_batch_blocks=2)                     # two batches
    d_out, k = dc.compress(d_in, x9=True)
    blob = d_out[:k].cpu().numpy()
    assert blob.tobytes() == oracle().compress_file_x9(data)
    offs, end = api.record_offsets(blob, 4, 16)
    assert end == k
    d_file = torch.zeros(k + 64, dtype=torch.uint8, device="cuda")
    d_file[:k] = d_out[:k]
    d_dec = torch.empty(4 << 20, dtype=torch.uint8, device="cuda")
    lens = dc.decompress(d_file, offs, d_dec, x9=True)
    assert int(lens.sum()) == len(data)
    assert torch.equal(d_dec[: len(data)], d_in)
    dc.close()


@pytest.mark.gpu
def test_gpu_x9_cli_streams_the_format_both_ways(tmp_path):
    """`-c fcx9` of the CLI: the multi-piece pipeline (2-block pieces) writes the same file as the one-call API and
    reads it back exactly, including a block the reference's format cannot round-trip."""
    import subprocess

    from my_compress_b200 import api, build

    build.build_cli()
    data = np.concatenate([datagen.gen("mixed", (2 << 20) + 999, seed=4), np.full(300000, 65, np.uint8)])
    src, out, back = tmp_path / "in.bin", tmp_path / "out.fcx9", tmp_path / "back.bin"
    src.write_bytes(data.tobytes())
    subprocess.check_call([build.CLI_SMALL, "-i", str(src), "-o", str(out), "-c", "fcx9"])
    assert out.read_bytes() == api.compress_x9(data)
    subprocess.check_call([build.CLI_SMALL, "-i", str(out), "-o", str(back)])
    assert back.read_bytes() == data.tobytes()
