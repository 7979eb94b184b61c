"""ctypes loaders for the checker libraries (oracle port + reference build).

Test infrastructure only: the product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libfcx_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmcref.so")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "my_compress_ref")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _p8(a):
    return a.ctypes.data_as(u8p)


def _p32(a):
    return a.ctypes.data_as(u32p)


def _as_u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def build_oracle():
    src = os.path.join(ORACLE_DIR, "fcx_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libfcx_oracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class _Codec:
    """Uniform python face over either library (prefix 'fcxo_' or 'ref_')."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.px = prefix
        f = self._f
        f("sunday_search", C.c_int32, [u8p, C.c_int32, u8p, C.c_int32])
        f("match_all", None, [u8p, C.c_uint32, u32p, u32p])
        f("lz77_parse", C.c_uint32, [u8p, C.c_uint32, u32p, u32p, u8p])
        f("huffman_section", C.c_uint32, [u8p, C.c_uint32, u8p])
        f("golomb_encode", C.c_uint32, [u32p, C.c_uint32, u32p])
        f("combine_bits", C.c_uint32, [u32p, C.c_uint32, C.c_uint32, u8p])
        f("compress_block", C.c_uint32, [u8p, C.c_uint32, u8p])
        f("decompress_block", C.c_uint32, [u8p, C.c_uint32, u8p, C.c_uint32])
        f("compress_file", C.c_uint64, [u8p, C.c_uint64, u8p])
        # LZ78 / FCX8 (scope row f-3): tokenizer, block payload, block decoder
        f("lz78_tokenize", C.c_uint32, [u8p, C.c_uint32, u32p, u8p])
        f("compress_block_lz78", C.c_uint32, [u8p, C.c_uint32, u8p])
        f("decompress_block_lz78", C.c_uint32, [u8p, C.c_uint32, u8p, C.c_uint32])

    def _f(self, name, res, args):
        fn = getattr(self.lib, self.px + name)
        fn.restype = res
        fn.argtypes = args
        setattr(self, "_" + name, fn)

    # ---- wrappers -------------------------------------------------
    def sunday_search(self, main, sub):
        m = _as_u8(main + b"\0\0")
        s = _as_u8(sub)
        return self._sunday_search(_p8(m), len(main), _p8(s), len(sub))

    def match_all(self, data):
        x = _as_u8(data)
        n = len(x)
        p = np.zeros(max(n, 1), np.uint32)
        l = np.zeros(max(n, 1), np.uint32)
        self._match_all(_p8(x), n, _p32(p), _p32(l))
        return p[:n], l[:n]

    def lz77_parse(self, data):
        x = _as_u8(data)
        n = len(x)
        p = np.zeros(max(n, 1), np.uint32)
        l = np.zeros(max(n, 1), np.uint32)
        c = np.zeros(max(n, 1), np.uint8)
        T = self._lz77_parse(_p8(x), n, _p32(p), _p32(l), _p8(c))
        return p[:T].copy(), l[:T].copy(), c[:T].copy()

    def huffman_section(self, data):
        x = _as_u8(data)
        out = np.zeros(len(x) + len(x) // 4 + 1024, np.uint8)
        k = self._huffman_section(_p8(x), len(x), _p8(out))
        return out[:k].tobytes()

    def golomb_encode(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.uint32)
        words = np.zeros(int(v.sum() // 4 + 3 * len(v)) // 32 + 8, np.uint32)
        g = self._golomb_encode(_p32(v), len(v), _p32(words))
        return words[:g].copy()

    def combine_bits(self, vals, bits=11):
        v = np.ascontiguousarray(vals, dtype=np.uint32)
        out = np.zeros((bits * len(v)) // 8 + 8, np.uint8)
        k = self._combine_bits(_p32(v), len(v), bits, _p8(out))
        return out[:k].tobytes()

    def compress_block(self, data):
        x = _as_u8(data)
        # the reference's caller hands in zero-initialised 2 MiB buffers (mc:4073,4088)
        xin = np.zeros(max(2 << 20, len(x)), np.uint8)
        xin[: len(x)] = x
        out = np.zeros(2 << 20, np.uint8)
        k = self._compress_block(_p8(xin), len(x), _p8(out))
        return out[:k].tobytes()

    def decompress_block(self, payload, cap=(1 << 20) + 8):
        x = _as_u8(payload)
        out = np.zeros(cap, np.uint8)
        k = self._decompress_block(_p8(x), len(x), _p8(out), cap)
        if k == 0xFFFFFFFF:
            return None
        return out[:k].tobytes()

    def lz78_tokenize(self, data):
        x = _as_u8(data)
        n = len(x)
        idx = np.zeros(max(n, 1), np.uint32)
        c = np.zeros(max(n, 1), np.uint8)
        T = self._lz78_tokenize(_p8(x if n else np.zeros(1, np.uint8)), n, _p32(idx), _p8(c))
        return idx[:T].copy(), c[:T].copy()

    def compress_block_lz78(self, data):
        x = _as_u8(data)
        xin = np.zeros(max(2 << 20, len(x)), np.uint8)       # zero-initialised 2 MiB buffers, as the reference's caller (mc:4073,4088)
        xin[: len(x)] = x
        out = np.zeros(4 << 20, np.uint8)
        k = self._compress_block_lz78(_p8(xin), len(x), _p8(out))
        return out[:k].tobytes()

    def decompress_block_lz78(self, payload, cap=(1 << 20) + 64):
        x = _as_u8(payload)
        out = np.zeros(cap, np.uint8)
        k = self._decompress_block_lz78(_p8(x), len(x), _p8(out), cap)
        if k == 0xFFFFFFFF:
            return None
        return out[:k].tobytes()

    def compress_file(self, data):
        x = _as_u8(data)
        nb = (len(x) + (1 << 20) - 1) >> 20
        out = np.zeros(10 + nb * (4 + (2 << 20)) + 64, np.uint8)
        xin = x if len(x) else np.zeros(1, np.uint8)
        k = self._compress_file(_p8(xin), len(x), _p8(out))
        return out[:k].tobytes()


class Oracle(_Codec):
    def __init__(self):
        lib = C.CDLL(build_oracle())
        super().__init__(lib, "fcxo_")
        lib.fcxo_huffman_tree.restype = C.c_uint32
        lib.fcxo_huffman_tree.argtypes = [u32p, u32p]
        lib.fcxo_huffman_codes.restype = C.c_uint32
        lib.fcxo_huffman_codes.argtypes = [C.c_uint32, u32p, u64p, u8p]
        lib.fcxo_serialize_block.restype = C.c_uint32
        lib.fcxo_serialize_block.argtypes = [u32p, u32p, u8p, C.c_uint32, u8p]
        lib.fcxo_decompress_file.restype = C.c_uint64
        lib.fcxo_decompress_file.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64]
        lib.fcxo_compress_bound.restype = C.c_uint64
        lib.fcxo_compress_bound.argtypes = [C.c_uint64]
        lib.fcxo_compress_blocks.restype = C.c_uint64
        lib.fcxo_compress_blocks.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, u8p]
        lib.fcxo_longest_match.restype = None
        lib.fcxo_longest_match.argtypes = [u8p, C.c_uint32, C.c_uint32, u32p, u32p]
        lib.fcxo_compress_file_lz78.restype = C.c_uint64
        lib.fcxo_compress_file_lz78.argtypes = [u8p, C.c_uint64, u8p]
        # FCX9 (scope row f-4): the opt-in format
        lib.fcxo_compress_block_x9.restype = C.c_uint32
        lib.fcxo_compress_block_x9.argtypes = [u8p, C.c_uint32, u8p]
        lib.fcxo_decompress_block_x9.restype = C.c_uint32
        lib.fcxo_decompress_block_x9.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32]
        lib.fcxo_compress_file_x9.restype = C.c_uint64
        lib.fcxo_compress_file_x9.argtypes = [u8p, C.c_uint64, u8p]

    def huffman_tree(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.uint32)
        assert len(w) == 256
        child = np.zeros(510, np.uint32)
        R = self.lib.fcxo_huffman_tree(_p32(w), _p32(child))
        return R, child

    def huffman_codes(self, weights):
        R, child = self.huffman_tree(weights)
        bits = np.zeros(256, np.uint64)
        lens = np.zeros(256, np.uint8)
        mx = self.lib.fcxo_huffman_codes(R, _p32(child), bits.ctypes.data_as(u64p), _p8(lens))
        return R, child, bits, lens, mx

    def longest_match_bf(self, data, i):
        x = _as_u8(data)
        p = C.c_uint32(0)
        l = C.c_uint32(0)
        self.lib.fcxo_longest_match(_p8(x), len(x), i, C.byref(p), C.byref(l))
        return p.value, l.value

    def serialize_block(self, p, l, c):
        p = np.ascontiguousarray(p, np.uint32)
        l = np.ascontiguousarray(l, np.uint32)
        c = np.ascontiguousarray(c, np.uint8)
        T = len(c)
        out = np.zeros(2 * T + 4096, np.uint8)
        pp = p if T else np.zeros(1, np.uint32)
        ll = l if T else np.zeros(1, np.uint32)
        cc = c if T else np.zeros(1, np.uint8)
        k = self.lib.fcxo_serialize_block(_p32(pp), _p32(ll), _p8(cc), T, _p8(out))
        return out[:k].tobytes()

    def decompress_file(self, blob, cap):
        x = _as_u8(blob)
        out = np.zeros(max(cap, 1), np.uint8)
        k = self.lib.fcxo_decompress_file(_p8(x), len(x), _p8(out), cap)
        if k == 0xFFFFFFFFFFFFFFFF:
            return None
        return out[:k].tobytes()

    def compress_bound(self, n):
        return int(self.lib.fcxo_compress_bound(n))

    def compress_block_x9(self, data):
        x = _as_u8(data)
        out = np.zeros(2 << 20, np.uint8)
        k = self.lib.fcxo_compress_block_x9(_p8(x), len(x), _p8(out))
        return out[:k].tobytes()

    def decompress_block_x9(self, payload, cap=(1 << 20) + 8):
        x = _as_u8(payload)
        out = np.zeros(cap, np.uint8)
        k = self.lib.fcxo_decompress_block_x9(_p8(x), len(x), _p8(out), cap)
        return None if k == 0xFFFFFFFF else out[:k].tobytes()

    def compress_file_x9(self, data):
        x = _as_u8(data)
        nb = (len(x) + (1 << 20) - 1) >> 20
        out = np.zeros(16 + nb * (4 + (2 << 20)) + 64, np.uint8)
        k = self.lib.fcxo_compress_file_x9(_p8(x if len(x) else np.zeros(1, np.uint8)), len(x), _p8(out))
        return out[:k].tobytes()

    def compress_file_lz78(self, data):
        """FCX8 container (the reference CLI's `-c <anything but lz77>`)."""
        x = _as_u8(data)
        nb = (len(x) + (1 << 20) - 1) >> 20
        out = np.zeros(10 + nb * (4 + (3 << 20)) + 64, np.uint8)
        k = self.lib.fcxo_compress_file_lz78(_p8(x if len(x) else np.zeros(1, np.uint8)), len(x), _p8(out))
        assert k != 0xFFFFFFFFFFFFFFFF
        return out[:k].tobytes()


class Reference(_Codec):
    """The unmodified reference functions (oracle/_ref/libmcref.so)."""

    def __init__(self):
        lib = C.CDLL(REF_SO)
        super().__init__(lib, "ref_")
        lib.ref_huffman_tree.restype = C.c_uint32
        lib.ref_huffman_tree.argtypes = [u32p, u32p]
        lib.ref_golomb_decode.restype = C.c_uint32
        lib.ref_golomb_decode.argtypes = [u32p, C.c_uint32, u32p, C.c_uint32]
        lib.ref_longest_match.restype = None
        lib.ref_longest_match.argtypes = [u8p, C.c_uint32, C.c_uint32, u32p, u32p, u8p]
        lib.ref_sizeof_head.restype = C.c_uint32

    def huffman_tree(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.uint32)
        nodes = np.zeros(511 * 4, np.uint32)
        R = self.lib.ref_huffman_tree(_p32(w), _p32(nodes))
        return R, nodes.reshape(511, 4)

    def golomb_decode(self, words, need):
        w = np.ascontiguousarray(words, dtype=np.uint32).copy()
        out = np.zeros(need + 8, np.uint32)
        k = self.lib.ref_golomb_decode(_p32(w), len(w), _p32(out), need)
        return out[:k].copy()


def have_ref():
    return os.path.exists(REF_SO)


_ORACLE = None
_REF = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def reference():
    global _REF
    if _REF is None:
        _REF = Reference()
    return _REF
