"""Seeded input buffers shared by the CPU and GPU parity tests."""
import numpy as np

from my_compress_b200 import datagen


def small_buffers(seed=1234, count=60, max_len=6000):
    """Buffers that shook out edge cases in the survey's differential runs (App. D):
    tiny alphabets, periodic runs, all-zero, lengths around the 3/4-byte and window limits."""
    rng = np.random.default_rng(seed)
    out = []
    for n in (0, 1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 257, 258, 259, 260, 261, 2046, 2047, 2048, 2049, 2050, 2307, 4097):
        out.append(("len%d_a3" % n, rng.integers(97, 100, size=n, dtype=np.uint8)))
    for k in range(count):
        n = int(rng.integers(1, max_len))
        kind = k % 8
        if kind == 0:
            b = rng.integers(0, 2, size=n, dtype=np.uint8) + 65
        elif kind == 1:
            b = rng.integers(0, 4, size=n, dtype=np.uint8) + 48
        elif kind == 2:
            b = rng.integers(0, 16, size=n, dtype=np.uint8)
        elif kind == 3:
            b = rng.integers(0, 256, size=n, dtype=np.uint8)
        elif kind == 4:
            b = datagen.periodic_block(n, seed, k)
        elif kind == 5:
            b = np.zeros(n, np.uint8)
        elif kind == 6:
            b = datagen.phrases_block(n, seed, k)
        else:
            b = datagen.zipf_block(n, seed, k)
            # splice long copies so some matches hit the 257 cap / 2047 distance
            if n > 3000:
                b[n - 600 : n - 100] = b[n - 600 - 2047 : n - 100 - 2047]
        out.append(("rnd%d_k%d_n%d" % (k, kind, n), np.ascontiguousarray(b, dtype=np.uint8)))
    return out


def stream_buffers(seed=99):
    """Byte streams for the Huffman section: 1/2/256 symbols, Fibonacci and equal weights."""
    rng = np.random.default_rng(seed)
    out = []
    out.append(("one_sym_1", np.array([7], np.uint8)))
    out.append(("one_sym_many", np.full(1000, 200, np.uint8)))
    out.append(("one_sym_zero", np.zeros(33, np.uint8)))
    out.append(("two_syms", np.array([1, 2, 1, 1], np.uint8)))
    out.append(("aaba", np.frombuffer(b"aaba", np.uint8)))
    out.append(("all256_equal", np.arange(256, dtype=np.uint8).repeat(3)))
    out.append(("all256_once", rng.permutation(256).astype(np.uint8)))
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    out.append(("fibonacci", rng.permutation(np.repeat(np.arange(len(fib), dtype=np.uint8) + 10, fib))))
    out.append(("pow2", rng.permutation(np.repeat(np.arange(12, dtype=np.uint8), [1 << k for k in range(12)]))))
    for k in range(40):
        n = int(rng.integers(1, 5000))
        a = int(rng.integers(1, 257))
        if k % 3 == 0:
            s = rng.integers(0, a, size=n)
        elif k % 3 == 1:
            s = np.minimum(rng.geometric(0.3, size=n) - 1, a - 1)
        else:
            s = np.where(rng.random(n) < 0.9, 5, rng.integers(0, a, size=n))
        out.append(("s%d_a%d_n%d" % (k, a, n), s.astype(np.uint8)))
    for n in (31, 32, 33, 63, 64, 65, 4095, 4096, 4097):
        out.append(("bits%d" % n, rng.integers(0, 2, size=n, dtype=np.uint8)))  # 1-bit codes: W boundaries
    return out
