"""Block-range sharding of one input over N ranks (one process per GPU) and container assembly.

Blocks are independent (my_compress.cpp's block loop, mc:4090-4122, shares no state between
iterations), so the N>1 path has no data-path collective: rank r compresses the contiguous
block range ``shard_ranges(n_blocks, world)[r]`` and the only thing that crosses ranks is the
finished byte string of each shard, concatenated in rank order behind the 10-byte header
(mc:101-109).  ``gather_container`` does that concatenation over ``torch.distributed`` for
callers that want the whole file on rank 0; the throughput bench never needs it.
"""
import numpy as np

BLOCK_BYTES = 1 << 20
HEADER_BYTES = 10


def n_blocks(nbytes):
    return (int(nbytes) + BLOCK_BYTES - 1) // BLOCK_BYTES


def shard_ranges(total_blocks, world):
    """Contiguous block ranges [b0, b1) per rank; the same split fcx_compress_lz77 uses per device."""
    return [(total_blocks * r // world, total_blocks * (r + 1) // world) for r in range(world)]


def shard_byte_range(nbytes, world, rank):
    b0, b1 = shard_ranges(n_blocks(nbytes), world)[rank]
    return min(nbytes, b0 * BLOCK_BYTES), min(nbytes, b1 * BLOCK_BYTES)


def file_header(total_bytes, total_blocks):
    """'F','C','X','7' | u32 bytes (wraps) | u16 blocks (wraps) -- mc:101-109, rewritten at mc:4128."""
    return b"FCX7" + (int(total_bytes) & 0xFFFFFFFF).to_bytes(4, "little") + (int(total_blocks) & 0xFFFF).to_bytes(2, "little")


def assemble_container(total_bytes, shard_records):
    """shard_records: per rank, the concatenated 'u32 size | payload' records of its block range."""
    return file_header(total_bytes, n_blocks(total_bytes)) + b"".join(shard_records)


def strip_header(blob):
    """Rank 0 compresses with FCX_WITH_HEADER; its records start after the header."""
    return blob[HEADER_BYTES:]


def gather_container(records, total_bytes, group=None, dst=0):
    """All ranks pass their shard's records (bytes); rank ``dst`` gets the whole container, others None.
    Works on any torch.distributed backend (gloo on CPU, nccl on GPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(records)], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(records):
        mine[: len(records)] = torch.from_numpy(np.frombuffer(records, dtype=np.uint8).copy()).to(dev)
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    if rank != dst:
        return None
    parts = [bufs[r][: sizes[r]].cpu().numpy().tobytes() for r in range(world)]
    return assemble_container(total_bytes, parts)
