#!/usr/bin/env python
"""bench.py -- compress throughput of the lz77 hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one pass of the whole device pipeline (match -> parse/split -> Huffman -> container)
over one batch of synthetic input.  N=1 workload: configs[1], the 64 MiB Zipf-16 stream; with
N ranks (torchrun, one rank per GPU) every rank compresses its own 64-block shard of an
N x 64 MiB corpus -- blocks share no state, so there is no data-path collective ("weak").

value  : input GB/s with the input already resident in HBM (CUDA events on the launch stream,
         L2 flushed between steps, max over ranks).
e2e    : the same metric through the host-buffer C-ABI call fcx_compress_lz77 (pinned host
         buffers, H2D + kernels + D2H inside the timed region).
--impl reference : the reference's own CPU code (oracle/_ref, built from /root/reference here;
         else the oracle port) on all host cores, bounded sample per step.
"""
import argparse
import ctypes as C
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
    "zipf-64MiB": ("zipf", 64 << 20),
    "randm-1MiB": ("randm", 1 << 20),
    "phrases-1GiB": ("phrases", 1 << 30),
    "uniform-1GiB": ("uniform", 1 << 30),
    "mixed-1GiB": ("mixed", 1 << 30),       # per-GPU share of configs[3] (8 GiB over 8 GPUs)
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def encode_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of k_encode_w per launch, from the committed ncu
    --set full capture of this workload (profiles/encode_traffic.json); None if there is none."""
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference / oracle on the host cores
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


def cpu_compress_blocks(fn, blocks, threads):
    """Compresses the given 1 MiB blocks with `threads` host threads; returns seconds."""
    from concurrent.futures import ThreadPoolExecutor

    bufs = []
    for blk in blocks:
        src = np.zeros(2 << 20, np.uint8)          # the reference's caller passes zeroed 2 MiB buffers (mc:4073,4088)
        src[: len(blk)] = blk
        bufs.append((src, len(blk), np.zeros(2 << 20, np.uint8)))
    t0 = time.perf_counter()
    if threads <= 1:
        for s, n, o in bufs:
            fn(s.ctypes.data, n, o.ctypes.data)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(lambda a: fn(a[0].ctypes.data, a[1], a[2].ctypes.data), bufs))
    return time.perf_counter() - t0


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
    blocks, idx = sample_blocks(family, per_gpu, args.seed, 0, threads)
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


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from my_compress_b200 import api

    api.lib()                               # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)   # plumbing only: barrier + max over ranks

    family, per_gpu = WORKLOADS[args.workload]
    first_block = rank * ((per_gpu + BLOCK - 1) // BLOCK)
    host = datagen.gen(family, per_gpu, seed=args.seed, first_block=first_block)
    pinned_in = torch.from_numpy(host).pin_memory()
    bound = api.compress_bound(per_gpu)
    pinned_out = torch.empty(bound, dtype=torch.uint8).pin_memory()
    d_in = pinned_in.to(dev, non_blocking=False)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    nblocks = (per_gpu + BLOCK - 1) // BLOCK
    dc = api.DeviceCompressor(local_rank, max_batch_blocks=min(nblocks, 1024))
    stream = torch.cuda.current_stream(dev)
    dc.use_stream(stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        dc.compress_async(d_in, d_out, with_header=(rank == 0), header_bytes=per_gpu * world, header_blocks=nblocks * world)

    for _ in range(max(3, args.warmup)):
        flush.fill_(1)
        step()
    out_bytes = dc.finish()

    # ---- timed region: K steps, CUDA events on the launch stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = dc.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)                       # untimed: evicts input/output/workspace lines from L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        evs.append((a, b))
    barrier()
    clocks = sampler.stop()
    launches = dc.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    out_bytes = dc.finish()
    stats = dc.stats()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = per_gpu * world / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage times (separate instrumented passes; explains `value`) ---------------------
    dc.enable_timing(True)
    for _ in range(args.steps):
        flush.fill_(1)
        step()
        dc.finish()
    stage = {k: v / args.steps for k, v in dc.stage_times().items()}
    dc.enable_timing(False)
    peak, peak_src = measured_peaks()
    enc_ms = stage.get("encode", 0.0)
    alg_bytes = stats["s_in"] + stats["s_out"]
    achieved = alg_bytes / (enc_ms * 1e-3) / 1e9 if enc_ms > 0 else 0.0
    stage_total = sum(stage.values())
    roofline = {"kernel": "k_encode_w", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": encode_traffic(args.workload), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": enc_ms,
                "share_of_step": enc_ms / stage_total if stage_total else None}
    match_ms = stage.get("match", 0.0)
    kernels = {k: {"ms": v, "share": v / stage_total if stage_total else None} for k, v in stage.items()}
    if match_ms > 0:
        kernels["match"]["byte_compares_per_s_vs_2047N"] = 2047.0 * per_gpu / (match_ms * 1e-3)
        kernels["match"]["input_GBps"] = per_gpu / (match_ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI call (H2D + kernels + D2H every step) -------------
    src_np, dst_np = pinned_in.numpy(), pinned_out.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        got = api.compress_into(src_np, dst_np, 1) if world == 1 else _e2e_shard(api, src_np, dst_np, local_rank)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = api.compress_into(src_np, dst_np, 1) if world == 1 else _e2e_shard(api, src_np, dst_np, local_rank)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = per_gpu * world / (float(te.item()) / e2e_steps) / 1e9

    if rank == 0:
        line = {
            "metric": "compress GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "family": family, "bytes_per_gpu": per_gpu, "blocks_per_gpu": nblocks,
                       "block_bytes": BLOCK, "window": 2047, "l2": "flushed between steps (512 MiB write, untimed)",
                       "partitioning": "contiguous block ranges per GPU, no collective", "seed": args.seed},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": per_gpu, "d2h_bytes_per_step": int(got),
                    "steps": e2e_steps, "api": "fcx_compress_lz77 (pinned host buffers)"},
            "roofline": roofline, "kernels": kernels,
            "output": {"container_bytes_per_gpu": out_bytes, "ratio": out_bytes / per_gpu, "tokens": stats["tokens"],
                       "matches": stats["matches"], "huffman_s_in": stats["s_in"], "huffman_s_out": stats["s_out"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, family, per_gpu)
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def _e2e_shard(api, src, dst, device_index):
    n = C.c_uint64(0)
    rc = api.lib().fcx_compress_lz77(src.ctypes.data, src.size, dst.ctypes.data, dst.size, C.byref(n), 1 << device_index)
    if rc:
        raise RuntimeError(api.lib().fcx_last_error().decode())
    return int(n.value)


def cpu_baseline(args, family, per_gpu):
    """Reference CPU path, single thread, bounded stride sample of the same workload (about 10-30 s)."""
    fn, kind = load_cpu_codec()
    count = {"zipf": 12, "uniform": 12, "randm": 1, "phrases": 64, "mixed": 12}.get(family, 8)
    blocks, idx = sample_blocks(family, per_gpu, args.seed, 0, count)
    nbytes = sum(len(b) for b in blocks)
    t = cpu_compress_blocks(fn, blocks, 1)
    return {"value": nbytes / t / 1e9, "unit": "GB/s", "cores": 1, "kind": kind, "seconds": t,
            "host_cores_available": os.cpu_count(),
            "sample": "%d of %d blocks of %s (stride sample, block indices %s...), single thread" % (
                len(blocks), (per_gpu + BLOCK - 1) // BLOCK, args.workload, idx[:4])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="zipf-64MiB", choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
