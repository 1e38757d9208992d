#!/bin/bash
# Builds libsag.so with AddressSanitizer on the host code of every translation unit (device code unchanged) and runs the CPU suite
# against it (no GPU needed): the C ABI's host paths -- marker parser / entropy decoders / emulated device decoder of jpeg.cu, the
# stream-K schedule, the EMD solver, crc32c.  The in-tree libsag.so is put back afterwards.
set -e
cd "$(dirname "$0")/.."
out=${TMPDIR:-/tmp}/sag_asan
mkdir -p "$out"
for f in spatialaudiogen_b200/csrc/*.cu; do
  b=$(basename "$f" .cu)
  nvcc -c "$f" -o "$out/$b.o" -O1 -g -std=c++17 -Xcompiler -fPIC -Xcompiler -fsanitize=address -Xcompiler -fno-omit-frame-pointer \
       -gencode arch=compute_100a,code=sm_100a &
done
wait
nvcc -shared -o "$out/libsag.so" "$out"/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fsanitize=address
cp spatialaudiogen_b200/libsag.so "$out/libsag_normal.so"
trap 'cp "$out/libsag_normal.so" spatialaudiogen_b200/libsag.so' EXIT
cp "$out/libsag.so" spatialaudiogen_b200/libsag.so
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests -q -m "not gpu" -p no:cacheprovider
