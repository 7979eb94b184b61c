"""CPU-side checks of the boundary: libfcx.so loads, exports everything include/fcx.h declares,
and refuses to compute without a CUDA device (there is no CPU fallback to fall into)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from my_compress_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build.build()
    return api.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "fcx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcx_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(L):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libfcx.so does not export %s" % n
    assert sorted(api.EXPORTS) == names


def test_compress_bound_covers_reference_worst_case(L):
    # worst observed payload: random data ~1.0163 n; hard bound 9/8 n + trees + counters
    for n in (0, 1, 1 << 20, (1 << 20) + 1, 64 << 20):
        nb = (n + (1 << 20) - 1) >> 20
        assert api.compress_bound(n) >= 10 + nb * (4 + 12 + 4 * 579) + n + n // 8


def test_argument_errors(L):
    out = np.zeros(64, np.uint8)
    n = C.c_uint64(0)
    assert L.fcx_compress_lz77(None, 5, out.ctypes.data, 64, C.byref(n), 1) == -1      # FCX_ERR_ARG (mc:2122 returns 0)
    assert b"null" in L.fcx_last_error()
    n32 = C.c_uint32(0)
    assert L.fcx_compress_block_lz77(out.ctypes.data, 0, out.ctypes.data, 64, C.byref(n32)) == -1
    assert L.fcx_compress_block_lz77(out.ctypes.data, (1 << 20) + 1, out.ctypes.data, 64, C.byref(n32)) == -1
    bad = np.frombuffer(b"NOPE000000000000", np.uint8)
    assert L.fcx_decompress_lz77(bad.ctypes.data, len(bad), out.ctypes.data, 64, C.byref(n), 1) == -5   # mc:4146-4152
    fcx6 = np.frombuffer(b"FCX6" + bytes(12), np.uint8)                 # '7' = lz77, '8' = lz78 (mc:4153-4164), '9' = the opt-in format
    assert L.fcx_decompress_lz77(fcx6.ctypes.data, len(fcx6), out.ctypes.data, 64, C.byref(n), 1) == -5
    # FCX9 header whose block count contradicts its byte count: rejected before anything is sized from it
    bad9 = np.frombuffer(b"FCX9" + (5).to_bytes(8, "little") + (0xFFFFFFFF).to_bytes(4, "little"), np.uint8)
    assert L.fcx_decompress_lz77(bad9.ctypes.data, len(bad9), out.ctypes.data, 64, C.byref(n), 1) in (-5, -2)


def test_context_batch_limit(L):
    # stream ids in the chunk descriptors and 32-bit arena offsets cap a device batch at 1024 blocks:
    # larger contexts are refused up front instead of corrupting the container
    ctx = C.c_void_p(None)
    assert L.fcx_ctx_create(C.byref(ctx), 0, 1025) == -1
    assert L.fcx_ctx_create(C.byref(ctx), 0, 0) == -1
    assert b"1..1024" in L.fcx_last_error()


def test_no_cpu_fallback_without_device(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert L.fcx_device_count() == 0
    x = np.arange(100, dtype=np.uint8)
    out = np.zeros(4096, np.uint8)
    n = C.c_uint64(0)
    assert L.fcx_compress_lz77(x.ctypes.data, 100, out.ctypes.data, 4096, C.byref(n), 1) == -2          # FCX_ERR_NO_DEVICE
    ctx = C.c_void_p(None)
    assert L.fcx_ctx_create(C.byref(ctx), 0, 4) == -2
    with pytest.raises(api.FcxError):
        api.compress(b"hello world")


def test_randm_generator_is_a_flat_permutation(L):
    # 随机数的生成.cpp:26-42: a permutation of [0,m); low bytes are exactly flat when 256 | m
    b = api.datagen_randm(1 << 16, 20261018)
    assert np.array_equal(np.bincount(b, minlength=256), np.full(256, 256))
    assert np.array_equal(b, api.datagen_randm(1 << 16, 20261018))
    assert not np.array_equal(b, api.datagen_randm(1 << 16, 7))


def test_pipeline_plan_covers_every_block_within_the_batch_limits(L):
    # the host path deals these batches to a device (csrc/fcx_host.cu): every block exactly once, no batch beyond the
    # 1024-block device limit, never more batches than the per-device event pool (256), short first and last batch
    sizes = (C.c_uint32 * 512)()
    for blocks in (0, 1, 2, 15, 16, 17, 40, 255, 256, 257, 511, 512, 513, 1000, 1024, 1025, 4096, 8192, 65535, 200000):
        k = L.fcx_debug_pipeline_plan(blocks, sizes, 512)
        plan = [int(sizes[i]) for i in range(min(k, 512))]
        assert k <= 256 and k == len(plan), (blocks, k)
        assert sum(plan) == blocks and all(0 < s <= 1024 for s in plan), (blocks, plan[:6], plan[-4:])
        if blocks >= 512:
            assert plan[0] == 8 and plan[-1] == 8 and max(plan) >= 128
        elif blocks > 16:
            assert plan[0] <= 16 and plan[-1] <= 16
