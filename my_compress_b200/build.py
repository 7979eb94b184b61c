"""Builds the C-ABI shared library (my_compress_b200/libfcx.so) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfcx.so")
SOURCES = ["fcx_api.cu", "fcx_host.cu", "fcx_match.cu", "fcx_parse.cu", "fcx_huff.cu", "fcx_decode.cu", "fcx_lz78.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-cudart", "shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fcx.h")]
    return any(os.path.getmtime(d) > t for d in deps)


CLI = os.path.join(HERE, "fcx_cli")


CLI_SMALL = os.path.join(HERE, "fcx_cli_smallpiece")      # test-only: 2-block pieces


def build_cli(force=False):
    """The flag-compatible CLI (my_compress_b200/cli/fcx_cli.cpp) linked against libfcx.so, and a test
    variant whose file pieces are 2 blocks so that small files walk the multi-piece pipeline."""
    src = os.path.join(HERE, "cli", "fcx_cli.cpp")
    for exe, defs in ((CLI, []), (CLI_SMALL, ["-DFCX_CLI_PIECE_BLOCKS=2"])):
        if not force and os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
            continue
        cmd = ["g++", "-O2", "-std=c++17", "-pthread"] + defs + [src, "-o", exe, "-L" + HERE, "-lfcx", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building " + os.path.basename(exe))
    return CLI


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libfcx.so")
    return LIB


# Test-only variant: a 96-word stage per warp, so that ordinary 8-bit streams overflow it and k_encode_w's
# direct (atomicOr from registers) path for over-dense sub-chunks is exercised by tests/test_gpu_parity.py.
LIB_SMALLSTAGE = os.path.join(HERE, "libfcx_smallw.so")


def build_test_variants(force=False):
    if not force and os.path.exists(LIB_SMALLSTAGE) and os.path.getmtime(LIB_SMALLSTAGE) >= os.path.getmtime(LIB):
        return LIB_SMALLSTAGE
    cmd = [NVCC] + FLAGS + ["-DFCX_ENCW_STAGE_WORDS=96"] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_SMALLSTAGE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libfcx_smallw.so")
    return LIB_SMALLSTAGE


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_cli(force=True)
    build_test_variants(force=True)
    print(LIB)
