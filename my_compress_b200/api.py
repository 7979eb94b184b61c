"""Python face of the C-ABI library (include/fcx.h), mirroring the reference's interface.

Names follow my_compress.cpp: ``my_compress_file_lz77`` (mc:2115) / ``my_decompress_file_lz77``
(mc:2255) work on one 1 MiB block payload; ``compress`` / ``decompress`` are the two branches
of ``main`` (mc:4073-4136, mc:4138-4202) on whole buffers.  Everything runs in the CUDA
kernels of ``libfcx.so``; there is no CPU path and nothing here touches ``oracle/``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("FCX_LIBFCX", os.path.join(_HERE, "libfcx.so"))   # override: tuning sweeps only
_lib = None

BLOCK_BYTES = 1 << 20
FCX_WITH_HEADER = 1
FCX_FORMAT_X9 = 2          # the opt-in FCX9 format (include/fcx.h): 64-bit header fields, named single-symbol sections
X9_HEADER_BYTES = 16

_u8p = C.POINTER(C.c_uint8)


class FcxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fcx error %d: %s" % (code, msg))
        self.code = code


class FcxStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("s_in", "s_out", "tokens", "matches", "container_bytes", "blocks")]


EXPORTS = [
    "fcx_last_error", "fcx_device_count", "fcx_compress_bound", "fcx_compress_block_lz77", "fcx_compress_lz77",
    "fcx_decompress_block_lz77", "fcx_decompress_lz77", "fcx_ctx_create", "fcx_ctx_destroy", "fcx_ctx_set_stream",
    "fcx_compress_lz77_device", "fcx_decompress_lz77_device", "fcx_ctx_finish", "fcx_ctx_launch_count", "fcx_ctx_stats", "fcx_ctx_enable_timing",
    "fcx_ctx_stage_times", "fcx_compress_bound_lz78", "fcx_compress_block_lz78", "fcx_compress_lz78", "fcx_decompress_block_lz78",
    "fcx_compress_records_lz77", "fcx_decompress_records_lz77", "fcx_host_alloc", "fcx_host_free",
    "fcx_compress_lz77_x9", "fcx_compress_records_lz77_x9", "fcx_compress_block_lz77_x9", "fcx_decompress_block_lz77_x9",
    "fcx_decompress_records_lz77_x9", "fcx_decompress_lz77_device_x9",
    "fcx_ctx_debug_guards", "fcx_debug_pipeline_plan", "fcx_probe_match", "fcx_probe_match_auto", "fcx_probe_streams", "fcx_probe_huffman_section", "fcx_datagen_randm",
]


def lib():
    """Loads libfcx.so (built in-tree by my_compress_b200/build.py); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FcxError(-2, "libfcx.so is not built: run `python -m my_compress_b200.build` (no CPU fallback exists)")
    L = C.CDLL(_LIB_PATH)
    L.fcx_last_error.restype = C.c_char_p
    L.fcx_device_count.restype = C.c_int
    L.fcx_compress_bound.restype = C.c_uint64
    L.fcx_compress_bound.argtypes = [C.c_uint64]
    L.fcx_compress_block_lz77.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.fcx_compress_lz77.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    L.fcx_compress_records_lz77.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    L.fcx_decompress_records_lz77.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    L.fcx_compress_bound_lz78.restype = C.c_uint64
    L.fcx_compress_bound_lz78.argtypes = [C.c_uint64]
    L.fcx_compress_block_lz78.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.fcx_compress_lz78.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    L.fcx_decompress_block_lz78.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.fcx_compress_lz77_x9.argtypes = L.fcx_compress_lz77.argtypes
    L.fcx_compress_records_lz77_x9.argtypes = L.fcx_compress_records_lz77.argtypes
    L.fcx_compress_block_lz77_x9.argtypes = L.fcx_compress_block_lz77.argtypes
    L.fcx_decompress_block_lz77_x9.argtypes = L.fcx_compress_block_lz77.argtypes
    L.fcx_decompress_records_lz77_x9.argtypes = L.fcx_decompress_records_lz77.argtypes
    L.fcx_host_alloc.restype = C.c_void_p
    L.fcx_host_alloc.argtypes = [C.c_uint64]
    L.fcx_host_free.restype = None
    L.fcx_host_free.argtypes = [C.c_void_p]
    L.fcx_decompress_block_lz77.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.fcx_decompress_lz77.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
    L.fcx_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint32]
    L.fcx_ctx_destroy.argtypes = [C.c_void_p]
    L.fcx_ctx_destroy.restype = None
    L.fcx_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.fcx_compress_lz77_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32,
                                           C.c_uint64, C.c_uint32]
    L.fcx_decompress_lz77_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.fcx_decompress_lz77_device_x9.argtypes = L.fcx_decompress_lz77_device.argtypes
    L.fcx_ctx_finish.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.fcx_ctx_launch_count.argtypes = [C.c_void_p]
    L.fcx_ctx_launch_count.restype = C.c_uint64
    L.fcx_ctx_stats.argtypes = [C.c_void_p, C.POINTER(FcxStats)]
    L.fcx_ctx_enable_timing.argtypes = [C.c_void_p, C.c_int]
    L.fcx_ctx_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.fcx_ctx_debug_guards.argtypes = [C.c_void_p, C.c_int]
    L.fcx_probe_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.fcx_probe_match_auto.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.fcx_probe_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fcx_probe_huffman_section.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_uint32)]
    L.fcx_debug_pipeline_plan.restype = C.c_uint32
    L.fcx_debug_pipeline_plan.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32]
    L.fcx_datagen_randm.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.fcx_datagen_randm.restype = None
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise FcxError(rc, lib().fcx_last_error().decode("utf-8", "replace"))


def _host_u8(buf):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def device_count():
    return int(lib().fcx_device_count())


def compress_bound(n):
    return int(lib().fcx_compress_bound(int(n)))


def device_mask(n_gpus):
    return (1 << int(n_gpus)) - 1


# ---- host-buffer API: what a user of the reference calls ---------------------------------

def my_compress_file_lz77(data):
    """One block (<= 1 MiB) -> payload bytes; same bytes as mc:2115 writes to pOut."""
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_compress_block_lz77(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


def my_decompress_file_lz77(payload):
    """One block payload -> the bytes mc:2255 writes to its FILE*."""
    x = _host_u8(payload)
    out = np.empty(BLOCK_BYTES + 8, np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_decompress_block_lz77(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


def compress_into(src, dst, n_gpus=1):
    """Whole buffer -> FCX7 container written into ``dst`` (numpy uint8, ideally pinned); returns byte count."""
    n = C.c_uint64(0)
    _check(lib().fcx_compress_lz77(src.ctypes.data, src.size, dst.ctypes.data, dst.size, C.byref(n), device_mask(n_gpus)))
    return int(n.value)


def compress(data, n_gpus=1):
    """``./my_compress -i in -o out -c lz77`` on a buffer: FCX7 container bytes."""
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    src = x if len(x) else np.zeros(1, np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_compress_lz77(src.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def compress_records(data, n_gpus=1):
    """The block loop alone (mc:4090-4122): records of ``data`` without the file header."""
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_compress_records_lz77(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def decompress_records(records, n_blocks, n_gpus=1):
    x = _host_u8(records)
    out = np.empty(max(1, n_blocks) * (BLOCK_BYTES + 8), np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_decompress_records_lz77(x.ctypes.data, len(x), int(n_blocks), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


# ---- FCX9: the opt-in format (64-bit header fields, named single-symbol sections; include/fcx.h) --------------

def compress_x9(data, n_gpus=1):
    """Whole buffer -> FCX9 container bytes (not readable by the reference; decodes back to exactly ``data``)."""
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    src = x if len(x) else np.zeros(1, np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_compress_lz77_x9(src.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def compress_records_x9(data, n_gpus=1):
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_compress_records_lz77_x9(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def decompress_records_x9(records, n_blocks, n_gpus=1):
    x = _host_u8(records)
    out = np.empty(max(1, n_blocks) * (BLOCK_BYTES + 8), np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_decompress_records_lz77_x9(x.ctypes.data, len(x), int(n_blocks), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def compress_block_x9(data):
    x = _host_u8(data)
    out = np.empty(compress_bound(len(x)), np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_compress_block_lz77_x9(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


def decompress_block_x9(payload):
    x = _host_u8(payload)
    out = np.empty(BLOCK_BYTES + 8, np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_decompress_block_lz77_x9(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


# ---- lz78 / FCX8 twins (mc:3127, mc:3478; "-c <anything but lz77>" of the reference CLI) -------------------

def my_compress_file_lz78(data):
    """One block (<= 1 MiB) -> FCX8 payload bytes; same bytes as mc:3127 writes to pOut."""
    x = _host_u8(data)
    out = np.empty(int(lib().fcx_compress_bound_lz78(len(x))), np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_compress_block_lz78(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


def my_decompress_file_lz78(payload):
    """One FCX8 block payload -> the bytes mc:3478 writes to its FILE* (a trailing 0x00 dropped, App. B2)."""
    x = _host_u8(payload)
    out = np.empty(BLOCK_BYTES + 8, np.uint8)
    n = C.c_uint32(0)
    _check(lib().fcx_decompress_block_lz78(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].tobytes()


def compress_lz78(data, n_gpus=1):
    """``./my_compress -i in -o out -c lz78`` on a buffer: FCX8 container bytes."""
    x = _host_u8(data)
    out = np.empty(int(lib().fcx_compress_bound_lz78(len(x))), np.uint8)
    src = x if len(x) else np.zeros(1, np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_compress_lz78(src.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


class PinnedBuffer:
    """Page-locked host memory from fcx_host_alloc as a numpy uint8 array (``.array``)."""

    def __init__(self, nbytes):
        self.ptr = lib().fcx_host_alloc(int(nbytes))
        if not self.ptr:
            raise FcxError(-3, lib().fcx_last_error().decode("utf-8", "replace"))
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _u8p), shape=(int(nbytes),))

    def close(self):
        if self.ptr:
            self.array = None
            lib().fcx_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decompress_into(src, dst, n_gpus=1):
    """FCX7 container (numpy uint8) -> decoded bytes written into ``dst`` (numpy uint8, ideally pinned); returns byte count."""
    n = C.c_uint64(0)
    _check(lib().fcx_decompress_lz77(src.ctypes.data, src.size, dst.ctypes.data, dst.size, C.byref(n), device_mask(n_gpus)))
    return int(n.value)


def decompress(blob, n_gpus=1):
    """``./my_compress -i in -o out`` on a buffer: what the reference's decoder writes."""
    x = _host_u8(blob)
    if len(x) < 10:
        raise FcxError(-5, "shorter than the 10-byte header")
    nb = int.from_bytes(x[8:10].tobytes(), "little")
    if x[3] == ord("9"):                                       # FCX9: u64 byte count, u32 block count
        if len(x) < X9_HEADER_BYTES:
            raise FcxError(-5, "shorter than the 16-byte FCX9 header")
        nb = min(int.from_bytes(x[12:16].tobytes(), "little"), len(x) // 4)
    out = np.empty(max(1, nb) * (BLOCK_BYTES + 8), np.uint8)
    n = C.c_uint64(0)
    _check(lib().fcx_decompress_lz77(x.ctypes.data, len(x), out.ctypes.data, len(out), C.byref(n), device_mask(n_gpus)))
    return out[: n.value].tobytes()


def record_offsets(blob, n_blocks, start=0):
    """Byte offsets of ``n_blocks`` consecutive 'u32 size | payload' records in ``blob`` (mc:4162-4174) and the end offset."""
    x = _host_u8(blob)
    offs = np.zeros(n_blocks, np.uint64)
    pos = int(start)
    for b in range(n_blocks):
        offs[b] = pos
        pos += 4 + int.from_bytes(x[pos : pos + 4].tobytes(), "little")
    return offs, pos


def datagen_randm(m, seed):
    out = np.empty(m, np.uint8)
    lib().fcx_datagen_randm(out.ctypes.data, m, seed)
    return out


# ---- device-resident API (torch tensors own the memory; kernels are ours) -----------------

class DeviceCompressor:
    """Compresses ``torch.uint8`` CUDA tensors in place in HBM on the current torch stream."""

    def __init__(self, device=0, max_batch_blocks=1024):
        import torch

        self.torch = torch
        self.device = int(device)
        self._ctx = C.c_void_p(None)
        _check(lib().fcx_ctx_create(C.byref(self._ctx), self.device, int(max_batch_blocks)))
        self.use_stream(torch.cuda.current_stream(self.device))

    def use_stream(self, stream):
        _check(lib().fcx_ctx_set_stream(self._ctx, C.c_void_p(stream.cuda_stream)))

    def close(self):
        if self._ctx:
            lib().fcx_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compress_async(self, d_in, d_out, with_header=True, header_bytes=None, header_blocks=None, n=None, x9=False):
        n = int(d_in.numel() if n is None else n)
        hb = n if header_bytes is None else int(header_bytes)
        hk = (n + BLOCK_BYTES - 1) // BLOCK_BYTES if header_blocks is None else int(header_blocks)
        _check(lib().fcx_compress_lz77_device(self._ctx, C.c_void_p(d_in.data_ptr()), n, C.c_void_p(d_out.data_ptr()),
                                              int(d_out.numel()), (FCX_WITH_HEADER if with_header else 0) | (FCX_FORMAT_X9 if x9 else 0), hb, hk & 0xFFFFFFFF))

    def finish(self):
        n = C.c_uint64(0)
        _check(lib().fcx_ctx_finish(self._ctx, C.byref(n)))
        return int(n.value)

    def compress(self, d_in, d_out=None, **kw):
        if d_out is None:
            d_out = self.torch.empty(compress_bound(d_in.numel()), dtype=self.torch.uint8, device=d_in.device)
        self.compress_async(d_in, d_out, **kw)
        return d_out, self.finish()

    def decompress(self, d_file, rec_off, d_out, x9=False):
        """Device-resident decode: block b of the records at ``rec_off`` (numpy uint64) -> d_out[b MiB:]; returns lens."""
        rec = np.ascontiguousarray(rec_off, dtype=np.uint64)
        lens = np.zeros(len(rec), np.uint32)
        fn = lib().fcx_decompress_lz77_device_x9 if x9 else lib().fcx_decompress_lz77_device
        _check(fn(self._ctx, C.c_void_p(d_file.data_ptr()), rec.ctypes.data, len(rec),
                                                C.c_void_p(d_out.data_ptr()), lens.ctypes.data))
        return lens

    def guards(self, check):
        """check=False paints the guard bands around every workspace array; check=True counts the overwritten ones."""
        rc = lib().fcx_ctx_debug_guards(self._ctx, 1 if check else 0)
        if rc < 0:
            _check(rc)
        return rc

    def launch_count(self):
        return int(lib().fcx_ctx_launch_count(self._ctx))

    def stats(self):
        s = FcxStats()
        _check(lib().fcx_ctx_stats(self._ctx, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in FcxStats._fields_}

    def enable_timing(self, on=True):
        _check(lib().fcx_ctx_enable_timing(self._ctx, 1 if on else 0))

    def stage_times(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        k = lib().fcx_ctx_stage_times(self._ctx, names, ms, 16)
        return {names[i].decode(): float(ms[i]) for i in range(k)}

    # ---- stage probes (tests) --------------------------------------------------------------
    def probe_match(self, d_in, auto=False):
        """auto=False: every position evaluated; auto=True: as the product path runs it (lazy tiles)."""
        t = self.torch
        n = d_in.numel()
        ln = t.empty(n, dtype=t.int16, device=d_in.device)
        ds = t.empty(n, dtype=t.int16, device=d_in.device)
        fn = lib().fcx_probe_match_auto if auto else lib().fcx_probe_match
        _check(fn(self._ctx, C.c_void_p(d_in.data_ptr()), n, C.c_void_p(ln.data_ptr()), C.c_void_p(ds.data_ptr())))
        return (ln.cpu().numpy().view(np.uint16).astype(np.uint32), ds.cpu().numpy().view(np.uint16).astype(np.uint32))

    def probe_streams(self, d_in):
        n = d_in.numel()
        T, Pc, G = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        bm = np.zeros(n // 8 + 64, np.uint8)
        cs = np.zeros(n + 64, np.uint8)
        pp = np.zeros(11 * (n // 4 + 1) // 8 + 64, np.uint8)
        gb = np.zeros(n // 8 + 1024, np.uint8)
        _check(lib().fcx_probe_streams(self._ctx, C.c_void_p(d_in.data_ptr()), n, C.byref(T), C.byref(Pc), C.byref(G),
                                       bm.ctypes.data, cs.ctypes.data, pp.ctypes.data, gb.ctypes.data))
        T, Pc, G = T.value, Pc.value, G.value
        return {"T": T, "Pc": Pc, "G": G, "bitmap": bm[: (T + 7) // 8].tobytes(), "c": cs[:T].tobytes(),
                "p": pp[: (11 * Pc) // 8 + 1].tobytes(), "golomb": gb[: 4 * G].tobytes()}

    def probe_huffman_section(self, d_src):
        S = d_src.numel()
        out = np.zeros(S + S // 4 + 4096, np.uint8)
        n = C.c_uint32(0)
        ptr = d_src.data_ptr() if S else self.torch.zeros(16, dtype=self.torch.uint8, device="cuda:%d" % self.device).data_ptr()
        _check(lib().fcx_probe_huffman_section(self._ctx, C.c_void_p(ptr), S, out.ctypes.data, len(out), C.byref(n)))
        return out[: n.value].tobytes()
