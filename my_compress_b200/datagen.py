"""Synthetic corpora for the five BASELINE.json configs.

Every generator is block-seeded: 1 MiB block ``b`` of a corpus depends only on
``(family, seed, b)``, so a rank can materialise just its own shard and the CPU
checker can regenerate any block on its own.
"""
import numpy as np

BLOCK = 1 << 20

PHRASES = [
    b"the quick brown fox jumps over the lazy dog. ",
    b"pack my box with five dozen liquor jugs; ",
    b"lorem ipsum dolor sit amet, consectetur adipiscing elit, ",
    b"0123456789 ",
    b"B200 NVLink HBM3e sm_100a ",
]


def _rng(family_id, seed, block):
    return np.random.Generator(np.random.PCG64([family_id, seed, block]))


def randm_block(n, seed=0, block=0):
    """Low bytes of the permutation RandM() of the reference's generator program produces
    (随机数的生成.cpp:26-42: Fisher-Yates driven by libc rand(), time() pinned to the seed), restated in
    ``fcx_datagen_randm`` (csrc/fcx_api.cu; host code, runs without a GPU): an exactly flat histogram
    when n is a multiple of 256."""
    from . import api

    if n == 0:
        return np.zeros(0, np.uint8)
    return api.datagen_randm(int(n), (int(seed) * 1000003 + int(block)) & 0x7FFFFFFF)


def uniform_block(n, seed=0, block=0):
    return _rng(2, seed, block).integers(0, 256, size=n, dtype=np.uint8)


_ZIPF_P = np.arange(1, 17, dtype=np.float64) ** -1.2
_ZIPF_P /= _ZIPF_P.sum()
_ZIPF_CDF = np.cumsum(_ZIPF_P)


def zipf_block(n, seed=0, block=0):
    """i.i.d. bytes over 16 symbols 'a'..'p', P(k) ~ k^-1.2."""
    u = _rng(3, seed, block).random(n)
    k = np.searchsorted(_ZIPF_CDF, u, side="right").clip(0, 15)
    return (k + ord("a")).astype(np.uint8)


_PH_TABLE = np.frombuffer(b"".join(PHRASES), dtype=np.uint8)
_PH_LEN = np.array([len(p) for p in PHRASES], dtype=np.int64)
_PH_BASE = np.concatenate([[0], np.cumsum(_PH_LEN)[:-1]])


def phrases_block(n, seed=0, block=0):
    """Uniform random concatenation of five fixed phrases (11-57 bytes)."""
    if n == 0:
        return np.zeros(0, np.uint8)
    count = n // int(_PH_LEN.min()) + 2
    idx = _rng(4, seed, block).integers(0, len(PHRASES), size=count)
    lens = _PH_LEN[idx]
    starts = np.cumsum(lens) - lens
    total = int(starts[-1] + lens[-1])
    src = np.repeat(_PH_BASE[idx] - starts, lens) + np.arange(total, dtype=np.int64)
    return _PH_TABLE[src[:n]]


def zeros_block(n, seed=0, block=0):
    return np.zeros(n, np.uint8)


def byte_run_block(n, seed=0, block=0):
    """A single repeated non-zero byte: every candidate matches to the cap (l=257 chain)."""
    return np.full(n, 0x41 + (seed + block) % 20, np.uint8)


def periodic_block(n, seed=0, block=0):
    r = _rng(5, seed, block)
    period = int(r.integers(2, 40))
    pat = r.integers(0, 256, size=period, dtype=np.uint8)
    return np.resize(pat, n)


FAMILIES = {
    "randm": randm_block,
    "uniform": uniform_block,
    "zipf": zipf_block,
    "phrases": phrases_block,
    "zeros": zeros_block,
    "byterun": byte_run_block,
    "periodic": periodic_block,
}
MIXED_ORDER = ("uniform", "zipf", "phrases")
PHRASES_VARIANT_BLOCK = 511      # one block of a single repeated byte (l = 257 chain, p = 2047 ties)


def block_family(family, block):
    if family == "mixed":
        return MIXED_ORDER[block % 3]
    if family == "phrases+run":                         # config 3 as the survey words it
        return "byterun" if block == PHRASES_VARIANT_BLOCK else "phrases"
    return family


def gen_block(family, n, seed=0, block=0):
    return FAMILIES[block_family(family, block)](n, seed, block)


def gen(family, nbytes, seed=0, first_block=0, out=None):
    """``nbytes`` of corpus ``family`` starting at 1 MiB block ``first_block`` (into ``out`` when given)."""
    if out is None:
        out = np.empty(nbytes, np.uint8)
    off = 0
    b = first_block
    while off < nbytes:
        n = min(BLOCK, nbytes - off)
        out[off : off + n] = gen_block(family, n, seed, b)
        off += n
        b += 1
    return out


CONFIGS = {
    # name -> (family, bytes)   BASELINE.json configs[0..4]
    "randm-1MiB": ("randm", 1 << 20),
    "zipf-64MiB": ("zipf", 64 << 20),
    "phrases-1GiB": ("phrases+run", 1 << 30),
    "mixed-8GiB": ("mixed", 8 << 30),
    "uniform-1GiB": ("uniform", 1 << 30),
}
