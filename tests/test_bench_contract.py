"""bench.py contract checks that need no GPU: the reference arm runs on the host cores and prints one
JSON line with the keys the driver reads; rank != 0 of a multi-rank launch prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({}, "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "randm-1MiB")
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "compress GB/s" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["config"]["workload"] == "randm-1MiB"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data"):
        assert k in line


def test_reference_arm_is_rank0_only():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""
