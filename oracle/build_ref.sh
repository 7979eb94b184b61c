#!/usr/bin/env bash
# Build the UNMODIFIED reference as a checker: oracle/_ref/libmcref.so (+ the CLI binary).
# TEST INFRASTRUCTURE ONLY.  Sources are compiled where they lie under /root/reference;
# only binaries land in oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#   library : lines 1..3710 of my_compress.cpp (all codec functions; main() excluded)
#             wrapped by oracle/ref_shim.cpp.  No patch needed.
#   CLI     : whole file with the one-token fix at mc:3732 (abs(uInt32) is ambiguous on
#             g++ 13; the line is scratch code in main(), not on the hot path).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${FCX_REFERENCE:-/root/reference}/my_compress.cpp"
OUT="$HERE/_ref"
if [ ! -f "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
head -n 3710 "$REF" > "$TMP/ref_body.inc"
g++ -O2 -w -fPIC -shared -std=gnu++17 -DREF_BODY="\"$TMP/ref_body.inc\"" \
    "$HERE/ref_shim.cpp" -o "$OUT/libmcref.so"
sed '3732s/abs(uib - uia)/abs((int)(uib - uia))/' "$REF" > "$TMP/my_compress_cli.cpp"
g++ -O2 -w -std=gnu++17 "$TMP/my_compress_cli.cpp" -o "$OUT/my_compress_ref"
echo "build_ref: built $OUT/libmcref.so and $OUT/my_compress_ref"
