#!/usr/bin/env python
"""Human-readable summary of a bench.py JSON line:  python tools/bench_summary.py gpurun_out/x.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def show(name, r):
    print("%-14s value %7.2f GB/s  %7.2f ms  e2e %6.2f  roof %.3f (enc %.3f ms)" % (
        name, r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"], r["roofline"]["launch_ms"]))
    k = r["kernels"]
    print("   stages:", {a: (round(b["ms"], 3) if isinstance(b, dict) else round(b, 3)) for a, b in k.items()})
    dec = r.get("decode") or {}
    print("   decode: dev %s GB/s (%s ms)  e2e %s  roundtrip %s %s" % (
        round(dec.get("value", 0), 2), round(dec.get("ms_per_step", 0), 2), round((dec.get("e2e") or {}).get("value", 0), 2),
        dec.get("roundtrip"), dec.get("error", "")))
    p = r.get("parity") or {}
    print("   parity:", p.get("status"), "blocks", p.get("blocks_checked_all_ranks"), "ranks", p.get("ranks"), p.get("error", ""))


show(d["config"]["workload"] + " N=%d" % d["n_gpus"], d)
for k, v in (d.get("configs") or {}).items():
    if "error" in v:
        print(k, v)
    else:
        show(k, v)
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"])
print("clocks", d["clocks"], "launches", d["gpu_launches"])
