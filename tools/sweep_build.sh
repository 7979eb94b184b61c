#!/usr/bin/env bash
# builds tuning variants of libfcx.so into gpurun_out/variants/ :  tools/sweep_build.sh NAME "-DFCX_X=.." ...
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p build/variants
name="$1"; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --shared -cudart shared "$@" \
  my_compress_b200/csrc/fcx_api.cu my_compress_b200/csrc/fcx_host.cu my_compress_b200/csrc/fcx_match.cu my_compress_b200/csrc/fcx_parse.cu \
  my_compress_b200/csrc/fcx_huff.cu my_compress_b200/csrc/fcx_decode.cu my_compress_b200/csrc/fcx_lz78.cu -o "build/variants/libfcx_${name}.so"
echo "built build/variants/libfcx_${name}.so"
