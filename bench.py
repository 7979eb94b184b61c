#!/usr/bin/env python
"""bench.py -- compress throughput of the lz77 hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one pass of the whole device pipeline (match -> parse/split -> Huffman -> container) over one
batch of synthetic input.  Default workload for every N: ``mixed-1GiB`` per GPU, i.e. rank r compresses
blocks [1024 r, 1024 (r+1)) of BASELINE configs[3] (the mixed corpus: block b is uniform / Zipf-16 /
phrases by b mod 3), so N = 8 IS the 8 GiB block-sharded config and N = 1 is its 1 GiB single-GPU share;
the file header written by rank 0 describes the whole N GiB file (its u32 byte count wraps to 0 at 8 GiB
as the reference's does).  Blocks share no state, so there is no data-path collective ("weak").

value   : input GB/s with the input already resident in HBM (CUDA events on the launch stream, L2 flushed
          between steps, max over ranks).
e2e     : the same metric through the host-buffer C-ABI call fcx_compress_lz77 (pinned host buffers,
          H2D + kernels + D2H inside the timed region).
roofline: k_encode_w, algorithmic bytes S_in + S_out over its CUDA-event time, against MEASURED_PEAKS.json.
decode  : fcx_decompress_lz77 on the container just produced, device-resident and end to end.
parity  : untimed, after the timed region: every rank compares >= 16 stride-sampled block records (rank 0
          also the file header) of the bytes it timed with the reference's CPU code (oracle/_ref, else the port).
configs : (N = 1 only) the other single-GPU BASELINE configs, each measured the same way.
--impl reference : the reference's own CPU code (oracle/_ref, built from /root/reference here; else the
          oracle port) on all host cores, bounded sample of the same workload per step.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from my_compress_b200 import datagen  # noqa: E402

BLOCK = 1 << 20
WORKLOADS = {
    # name -> (family, bytes per GPU)
    "mixed-1GiB": ("mixed", 1 << 30),            # per-GPU share of configs[3] (8 GiB over 8 GPUs)
    "zipf-64MiB": ("zipf", 64 << 20),            # configs[1]
    "randm-1MiB": ("randm", 1 << 20),            # configs[0]
    "phrases-1GiB": ("phrases+run", 1 << 30),    # configs[2] (+ the single-repeated-byte variant block)
    "uniform-1GiB": ("uniform", 1 << 30),        # configs[4]
}
DEFAULT_WORKLOAD = "mixed-1GiB"
EXTRA_CONFIGS = ("uniform-1GiB", "phrases-1GiB", "zipf-64MiB")     # reported under "configs" at N = 1
PARITY_BLOCKS = 16


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def encode_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of k_encode_w per launch, from the committed ncu
    --set full capture of this workload (profiles/encode_traffic.json, written by tools/ncu_traffic.py);
    None if there is none."""
    path = os.path.join(ROOT, "profiles", "encode_traffic.json")
    try:
        return json.load(open(path)).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        # "under load": the samples taken while the kernels ran draw visibly more power than the idle ones
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * (min(pw) + max(pw))] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference / oracle on the host cores (the checker; never on the product path)
# ------------------------------------------------------------------------------------------
def load_cpu_codec():
    ref = os.path.join(ROOT, "oracle", "_ref", "libmcref.so")
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        fn, kind = lib.ref_compress_block, "reference"
    else:
        so = os.path.join(ROOT, "oracle", "libfcx_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libfcx_oracle.so"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(so)
        fn, kind = lib.fcxo_compress_block, "port"
    fn.restype = C.c_uint32
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    return fn, kind


def cpu_compress_blocks(fn, blocks, threads, keep=False):
    """Compresses the given 1 MiB blocks with `threads` host threads; returns seconds (and the payloads)."""
    from concurrent.futures import ThreadPoolExecutor

    bufs = []
    for blk in blocks:
        src = np.zeros(2 << 20, np.uint8)          # the reference's caller passes zeroed 2 MiB buffers (mc:4073,4088)
        src[: len(blk)] = blk
        bufs.append((src, len(blk), np.zeros(2 << 20, np.uint8)))
    t0 = time.perf_counter()
    if threads <= 1:
        sizes = [fn(s.ctypes.data, n, o.ctypes.data) for s, n, o in bufs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            sizes = list(ex.map(lambda a: fn(a[0].ctypes.data, a[1], a[2].ctypes.data), bufs))
    dt = time.perf_counter() - t0
    if keep:
        return dt, [o[:k].tobytes() for (_, _, o), k in zip(bufs, sizes)]
    return dt


def sample_blocks(family, nbytes, seed, first_block, count):
    nb = (nbytes + BLOCK - 1) // BLOCK
    count = max(1, min(count, nb))
    idx = [int(i * nb / count) for i in range(count)]
    out = []
    for b in idx:
        n = min(BLOCK, nbytes - b * BLOCK)
        out.append(datagen.gen_block(family, n, seed, first_block + b))
    return out, idx


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    fn, kind = load_cpu_codec()
    family, per_gpu = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    count = threads if family != "mixed" else max(3, threads - threads % 3)     # the three families in equal parts
    blocks, idx = sample_blocks(family, per_gpu, args.seed, 0, count)
    nbytes = sum(len(b) for b in blocks)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_compress_blocks(fn, blocks[: max(1, threads // 4)], threads)
    times = [cpu_compress_blocks(fn, blocks, threads) for _ in range(args.steps)]
    t = sum(times) / len(times)
    val = nbytes / t / 1e9
    line = {
        "impl": "reference", "metric": "compress GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "family": family, "sample_bytes_per_step": nbytes, "host_threads": threads},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind,
                         "sample": "%d blocks of %s (block indices %s...), %d host threads, per step" % (len(blocks), args.workload, idx[:4], threads)},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline(args, workload):
    """Reference CPU path, single thread, bounded stride sample of the same workload (about 10-30 s)."""
    fn, kind = load_cpu_codec()
    family, per_gpu = WORKLOADS[workload]
    count = {"zipf": 12, "uniform": 12, "randm": 1, "phrases": 64, "phrases+run": 64, "mixed": 12}.get(family, 8)
    blocks, idx = sample_blocks(family, per_gpu, args.seed, 0, count)
    nbytes = sum(len(b) for b in blocks)
    t = cpu_compress_blocks(fn, blocks, 1)
    return {"value": nbytes / t / 1e9, "unit": "GB/s", "cores": 1, "kind": kind, "seconds": t,
            "host_cores_available": os.cpu_count(),
            "sample": "%d of %d blocks of %s (stride sample, block indices %s...), single thread" % (
                len(blocks), (per_gpu + BLOCK - 1) // BLOCK, workload, idx[:4])}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class Rig:
    """Per-process state shared by the workloads measured in one run."""

    def __init__(self, args, rank, world, local_rank):
        import torch

        self.torch = torch
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.distributed = world > 1
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.stream = torch.cuda.current_stream(self.dev)

    def barrier(self):
        import torch.distributed as dist

        self.torch.cuda.synchronize(self.dev)
        if self.distributed:
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        import torch.distributed as dist

        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        import torch.distributed as dist

        t = self.torch.tensor(xs, dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def parity_check(rig, api, family, per_gpu, first_block, host_in, blob, nblocks, has_header, hdr_bytes, hdr_blocks):
    """Compares stride-sampled block records of `blob` (numpy uint8: what the timed steps produced) with the
    reference's CPU code run on the same input blocks.  Untimed."""
    fn, kind = load_cpu_codec()
    start = 10 if has_header else 0
    offs, end = api.record_offsets(blob, nblocks, start)
    count = min(PARITY_BLOCKS, nblocks)
    idx = sorted(set(int(i * nblocks / count) for i in range(count)) | {nblocks - 1})
    blocks = [host_in[b * BLOCK : min(per_gpu, (b + 1) * BLOCK)] for b in idx]
    threads = max(1, min(len(idx), (os.cpu_count() or 1) // max(1, min(rig.world, 8)) or 1, 16))
    _, payloads = cpu_compress_blocks(fn, blocks, threads, keep=True)
    mism = []
    for b, want in zip(idx, payloads):
        o = int(offs[b])
        nxt = int(offs[b + 1]) if b + 1 < nblocks else end
        got = blob[o:nxt].tobytes()
        rec = len(want).to_bytes(4, "little") + want
        if hashlib.sha256(got).digest() != hashlib.sha256(rec).digest():
            mism.append(first_block + b)
    header_ok = None
    if has_header:
        want = b"FCX7" + (hdr_bytes & 0xFFFFFFFF).to_bytes(4, "little") + (hdr_blocks & 0xFFFF).to_bytes(2, "little")
        header_ok = blob[:10].tobytes() == want
    ok = not mism and header_ok is not False and end == len(blob)
    return {"status": "ok" if ok else "MISMATCH", "checker": kind, "blocks_checked": len(idx), "mismatched_blocks": mism,
            "header_ok": header_ok, "records_span_container": end == len(blob),
            "sample": "records of blocks %s... of this rank's shard, sha256 vs the CPU %s" % (idx[:4], kind)}


def measure(rig, api, workload, steps, warmup, want_cpu_baseline):
    """One workload on this rank's GPU: device-resident value, stage times + roofline, e2e, decode, parity."""
    torch = rig.torch
    args, rank, world, dev, stream = rig.args, rig.rank, rig.world, rig.dev, rig.stream
    family, per_gpu = WORKLOADS[workload]
    nblocks = (per_gpu + BLOCK - 1) // BLOCK
    first_block = rank * nblocks
    pinned_in = torch.empty(per_gpu, dtype=torch.uint8, pin_memory=True)
    host_in = pinned_in.numpy()
    datagen.gen(family, per_gpu, seed=args.seed, first_block=first_block, out=host_in)
    bound = api.compress_bound(per_gpu)
    pinned_out = torch.empty(bound, dtype=torch.uint8, pin_memory=True)
    d_in = pinned_in.to(dev, non_blocking=False)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    dc = api.DeviceCompressor(rig.local_rank, max_batch_blocks=min(nblocks, 1024))
    dc.use_stream(stream)
    hdr_bytes, hdr_blocks = per_gpu * world, nblocks * world

    def step():
        dc.compress_async(d_in, d_out, with_header=(rank == 0), header_bytes=hdr_bytes, header_blocks=hdr_blocks)

    for _ in range(max(3, warmup)):
        rig.flush.fill_(1)
        step()
    dc.finish()

    # ---- timed region: K steps, CUDA events on the launch stream, L2 flushed between steps ----
    sampler = ClockSampler(rig.local_rank)
    sampler.start()
    rig.barrier()
    launches0 = dc.launch_count()
    evs = []
    for _ in range(steps):
        rig.flush.fill_(1)                   # untimed: evicts input/output/workspace lines from L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        evs.append((a, b))
    rig.barrier()
    clocks = sampler.stop()
    launches = dc.launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    out_bytes = dc.finish()
    stats = dc.stats()
    ms_per_step = rig.max_over_ranks(total_ms) / steps
    value = per_gpu * world / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage times (separate instrumented passes; explains `value`) ---------------------
    dc.enable_timing(True)
    for _ in range(steps):
        rig.flush.fill_(1)
        step()
        dc.finish()
    stage = {k: v / steps for k, v in dc.stage_times().items()}
    dc.enable_timing(False)
    peak, peak_src = measured_peaks()
    enc_ms = stage.get("encode", 0.0)
    alg_bytes = stats["s_in"] + stats["s_out"]
    achieved = alg_bytes / (enc_ms * 1e-3) / 1e9 if enc_ms > 0 else 0.0
    stage_total = sum(stage.values())
    roofline = {"kernel": "k_encode_w", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": encode_traffic(workload), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": enc_ms,
                "share_of_step": enc_ms / stage_total if stage_total else None,
                "launches_per_step": (nblocks + 1023) // 1024}
    kernels = {k: {"ms": v, "share": v / stage_total if stage_total else None} for k, v in stage.items()}
    match_ms = stage.get("match", 0.0)
    if match_ms > 0:
        kernels["match"]["byte_compares_per_s_vs_2047N"] = 2047.0 * per_gpu / (match_ms * 1e-3)
        kernels["match"]["input_GBps"] = per_gpu / (match_ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI call (H2D + kernels + D2H every step) -------------
    src_np, dst_np = host_in, pinned_out.numpy()
    e2e_steps = max(3, min(steps, 10))

    def e2e_call():
        n = C.c_uint64(0)
        rc = api.lib().fcx_compress_lz77(src_np.ctypes.data, src_np.size, dst_np.ctypes.data, dst_np.size, C.byref(n), 1 << rig.local_rank)
        if rc:
            raise RuntimeError(api.lib().fcx_last_error().decode())
        return int(n.value)

    for _ in range(2):
        got = e2e_call()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = e2e_call()
    torch.cuda.synchronize(dev)
    e2e_s = rig.max_over_ranks(time.perf_counter() - t0)
    e2e_value = per_gpu * world / (e2e_s / e2e_steps) / 1e9
    e2e_blob = dst_np[:got].copy()

    # ---- the bytes the timed steps produced, on the host (untimed) --------------------------------
    blob = d_out[:out_bytes].cpu().numpy()
    hdr = 10 if rank == 0 else 0
    # the e2e call compresses the shard as a file of its own: same records behind its own header
    e2e_same = bool(len(e2e_blob) - 10 == len(blob) - hdr and np.array_equal(e2e_blob[10:], blob[hdr:]))

    res = {
        "value": value, "ms_per_step": ms_per_step, "gpu_launches": launches, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": per_gpu, "d2h_bytes_per_step": int(got),
                "steps": e2e_steps, "api": "fcx_compress_lz77 (pinned host buffers)", "records_equal_device_path": e2e_same},
        "roofline": roofline, "kernels": kernels,
        "output": {"container_bytes_per_gpu": out_bytes, "ratio": out_bytes / per_gpu, "tokens": stats["tokens"],
                   "matches": stats["matches"], "huffman_s_in": stats["s_in"], "huffman_s_out": stats["s_out"]},
    }

    # ---- decode: the container just produced, device-resident and through the host-buffer call ---------
    # (no collective inside the try: a rank that fails here must not leave the others waiting)
    dec = None
    dec_ms_local, dec_s_local = -1.0, -1.0
    try:
        offs, end = api.record_offsets(blob, nblocks, hdr)
        d_file = torch.zeros(len(blob) + 64, dtype=torch.uint8, device=dev)
        d_file[: len(blob)] = d_out[:out_bytes]
        d_dec = torch.empty(nblocks * BLOCK, dtype=torch.uint8, device=dev)
        dec_steps = max(2, min(steps, 5))
        for _ in range(2):
            lens = dc.decompress(d_file, offs, d_dec)
        dec_ms_local = 0.0
        for _ in range(dec_steps):
            rig.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            lens = dc.decompress(d_file, offs, d_dec)
            b.record(stream)
            torch.cuda.synchronize(dev)
            dec_ms_local += a.elapsed_time(b)
        dec_ms_local /= dec_steps
        full = bool(int(lens.sum()) == per_gpu)
        diff_blocks = 0
        if full:
            same = (d_dec[:per_gpu] == d_in)
            if not bool(same.all()):
                pad = torch.ones(nblocks * BLOCK, dtype=torch.bool, device=dev)
                pad[:per_gpu] = same
                diff_blocks = int((~pad.view(nblocks, BLOCK).all(dim=1)).sum().item())
        # end to end: records in pinned host memory -> decoded bytes in pinned host memory
        rec_host = torch.empty(len(blob) - hdr, dtype=torch.uint8, pin_memory=True)
        rec_host.numpy()[:] = blob[hdr:]
        dec_host = torch.empty(nblocks * (BLOCK + 8), dtype=torch.uint8, pin_memory=True)
        rec_np, dec_np = rec_host.numpy(), dec_host.numpy()

        def dec_call():
            n = C.c_uint64(0)
            rc = api.lib().fcx_decompress_records_lz77(rec_np.ctypes.data, rec_np.size, nblocks, dec_np.ctypes.data, dec_np.size,
                                                       C.byref(n), 1 << rig.local_rank)
            if rc:
                raise RuntimeError(api.lib().fcx_last_error().decode())
            return int(n.value)

        for _ in range(2):
            dgot = dec_call()
        t0 = time.perf_counter()
        for _ in range(dec_steps):
            dgot = dec_call()
        dec_s_local = (time.perf_counter() - t0) / dec_steps
        e2e_equal = bool(dgot == per_gpu and np.array_equal(dec_np[:dgot], d_dec[:per_gpu].cpu().numpy()))
        # a single-symbol Huffman stream loses its symbol in the reference's format: the
        # single-repeated-byte variant block of phrases-1GiB decodes to what the reference's decoder yields, not to the input
        expected_lossy = 1 if (family == "phrases+run" and first_block <= datagen.PHRASES_VARIANT_BLOCK < first_block + nblocks) else 0
        dec = {
            "unit": "GB/s (decoded bytes, device-resident, k_decode)", "steps": dec_steps,
            "algorithmic_bytes": int(len(blob) + per_gpu),
            "e2e": {"unit": "GB/s", "api": "fcx_decompress_records_lz77 (pinned host buffers)",
                    "h2d_bytes_per_step": int(len(blob) - hdr), "d2h_bytes_per_step": int(dgot), "equals_device_path": e2e_equal},
            "decoded_bytes_ok": full, "blocks_differing_from_input": diff_blocks, "blocks_expected_lossy_by_format": expected_lossy,
            "roundtrip": "ok" if (full and diff_blocks == expected_lossy and e2e_equal) else "MISMATCH",
        }
        del d_file, d_dec, rec_host, dec_host
    except Exception as e:                                   # the compress line must survive a decode problem
        dec = {"error": repr(e)[:300]}
    dec_ms = rig.max_over_ranks(dec_ms_local)
    dec_s = rig.max_over_ranks(dec_s_local)
    if "error" not in dec and dec_ms > 0 and dec_s > 0:
        dec["value"] = per_gpu * world / (dec_ms * 1e-3) / 1e9
        dec["ms_per_step"] = dec_ms
        dec["hbm_GBps"] = dec["algorithmic_bytes"] / (dec_ms * 1e-3) / 1e9
        dec["e2e"]["value"] = per_gpu * world / dec_s / 1e9
    res["decode"] = dec

    # ---- parity of the timed bytes against the reference's CPU code (untimed) -------------------
    try:
        par = parity_check(rig, api, family, per_gpu, first_block, host_in, blob, nblocks, rank == 0, hdr_bytes, hdr_blocks)
    except Exception as e:
        par = {"status": "ERROR", "error": repr(e)[:300], "blocks_checked": 0, "mismatched_blocks": []}
    bad = 0 if par["status"] == "ok" else 1
    tot = rig.sum_over_ranks([bad, par.get("blocks_checked", 0), 0 if e2e_same else 1])
    par["ranks"] = world
    par["ranks_failed"] = int(tot[0])
    par["blocks_checked_all_ranks"] = int(tot[1])
    par["e2e_records_differ_ranks"] = int(tot[2])
    if tot[0] or tot[2]:
        par["status"] = "MISMATCH"
    res["parity"] = par
    if want_cpu_baseline and rank == 0:
        res["cpu_baseline"] = cpu_baseline(args, workload)
    dc.close()
    del d_in, d_out, pinned_in, pinned_out
    torch.cuda.empty_cache()
    return res


def bind_near_gpu(index):
    """Pins this process to the host cores NVML reports as local to GPU `index`, so that its pinned buffers are
    allocated on that NUMA node (what numactl would do for a production caller; eight ranks that all stage through
    one socket share its memory and PCIe root).  Returns the number of cores, 0 if nothing was changed."""
    try:
        import pynvml

        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (ncpu + 63) // 64)
        cpus = [c for c in range(ncpu) if (words[c // 64] >> (c % 64)) & 1]
        if cpus and hasattr(os, "sched_setaffinity"):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_ours(args, rank, world, local_rank):
    near = bind_near_gpu(local_rank) if world > 1 else 0
    import torch
    import torch.distributed as dist

    from my_compress_b200 import api

    api.lib()                               # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    rig = Rig(args, rank, world, local_rank)
    if rig.distributed:
        dist.init_process_group("nccl", device_id=rig.dev)   # plumbing only: barrier + max over ranks

    family, per_gpu = WORKLOADS[args.workload]
    nblocks = (per_gpu + BLOCK - 1) // BLOCK
    main = measure(rig, api, args.workload, args.steps, args.warmup, world == 1 and not args.no_cpu_baseline)
    configs = {}
    if world == 1 and not args.no_configs and args.workload == DEFAULT_WORKLOAD:
        for w in EXTRA_CONFIGS:
            try:
                r = measure(rig, api, w, max(3, min(args.steps, 5)), args.warmup, False)
                configs[w] = {"value": r["value"], "unit": "GB/s", "ms_per_step": r["ms_per_step"], "e2e": r["e2e"],
                              "roofline": r["roofline"], "kernels": {k: v["ms"] for k, v in r["kernels"].items()},
                              "decode": r.get("decode"), "parity": r.get("parity"), "output": r["output"]}
            except Exception as e:
                configs[w] = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": "compress GB/s", "value": main["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "family": family, "bytes_per_gpu": per_gpu, "blocks_per_gpu": nblocks,
                       "total_bytes": per_gpu * world, "block_bytes": BLOCK, "window": 2047,
                       "l2": "flushed between steps (512 MiB write, untimed)",
                       "partitioning": "rank r holds blocks [%d r, %d (r+1)) of the %d-block file; contiguous block ranges, no collective" % (nblocks, nblocks, nblocks * world),
                       "host_cores_bound_near_gpu": near, "seed": args.seed},
            "clocks": main["clocks"], "gpu_launches": main["gpu_launches"],
            "e2e": main["e2e"], "roofline": main["roofline"], "kernels": main["kernels"], "output": main["output"],
            "decode": main.get("decode"), "parity": main.get("parity"),
        }
        if "cpu_baseline" in main:
            line["cpu_baseline"] = main["cpu_baseline"]
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if rig.distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other single-GPU configs (N = 1)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: launch N>1 with torchrun (one rank per GPU); see the module docstring")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
