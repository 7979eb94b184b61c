#!/usr/bin/env python
"""profiles/encode_traffic.json from ncu --set full reports of k_encode_w, one per workload:
    python tools/ncu_traffic.py mixed-1GiB=gpurun_out/x_mixed.ncu-rep uniform-1GiB=... [--source "text"]
dram__bytes_read.sum + dram__bytes_write.sum of the k_encode_w launch (the last one if several), per launch.
bench.py reads the file for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def kernel_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if "k_encode_w" in r[idx["Kernel Name"]]:
            def val(name):
                return float(r[idx[name]].replace(",", "")) * UNIT.get(units[idx[name]], 1)
            yield {"dram_read_bytes": int(val("dram__bytes_read.sum")), "dram_write_bytes": int(val("dram__bytes_write.sum")),
                   "duration_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0)}


def main():
    source = "ncu --set full --clock-control none -k regex:k_encode_w, tools/prof_run.py"
    args = [a for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    if "--source" in sys.argv:
        source = sys.argv[sys.argv.index("--source") + 1]
    path = os.path.join(ROOT, "profiles", "encode_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for a in args:
        name, rep = a.split("=", 1)
        rows = list(kernel_rows(rep))
        if not rows:
            print("no k_encode_w launch in", rep)
            continue
        r = rows[-1]
        data[name] = {"dram_bytes_per_launch": r["dram_read_bytes"] + r["dram_write_bytes"], "dram_read_bytes": r["dram_read_bytes"],
                      "dram_write_bytes": r["dram_write_bytes"], "launch_us_under_ncu": r["duration_us"], "source": source + " (" + os.path.basename(rep) + ")"}
        print(name, data[name])
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
