"""N>1 host-side path on CPU: two gloo ranks each take a contiguous block range, produce the
records of their shard (with the oracle standing in for the GPU, test-only), and rank 0
assembles the container -- which must equal the single-process oracle container byte for byte."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from my_compress_b200 import datagen, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nbytes, outdir):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _libs import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_byte_range(nbytes, world, rank)
    b0, _ = sharding.shard_ranges(sharding.n_blocks(nbytes), world)[rank]
    shard = datagen.gen("mixed", hi - lo, seed=9, first_block=b0)        # a rank materialises only its own blocks
    blob = oracle().compress_file(shard) if hi > lo else b"FCX7" + bytes(6)
    records = sharding.strip_header(blob)
    whole = sharding.gather_container(records, nbytes, dst=0)
    if rank == 0:
        with open(os.path.join(outdir, "sharded.fcx7"), "wb") as f:
            f.write(whole)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_container_equals_single_process(tmp_path):
    from _libs import oracle

    nbytes = (3 << 20) + 4321                       # 4 blocks, ragged tail: ranks get 2 + 2
    mp.spawn(_worker, args=(2, _free_port(), nbytes, str(tmp_path)), nprocs=2, join=True)
    want = oracle().compress_file(datagen.gen("mixed", nbytes, seed=9))
    got = open(os.path.join(str(tmp_path), "sharded.fcx7"), "rb").read()
    assert got == want


def test_shard_ranges_cover_everything_once():
    for blocks in (0, 1, 2, 7, 64, 8192):
        for world in (1, 2, 4, 8):
            r = sharding.shard_ranges(blocks, world)
            assert r[0][0] == 0 and r[-1][1] == blocks
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    # header fields wrap like the reference's uInt32 / uInt16 (mc:105-106): 8 GiB -> 0 bytes, 8192 blocks
    assert sharding.file_header(8 << 30, 8192) == b"FCX7" + bytes(4) + (8192).to_bytes(2, "little")
    assert sharding.file_header(5, 65536 + 3)[8:] == (3).to_bytes(2, "little")
