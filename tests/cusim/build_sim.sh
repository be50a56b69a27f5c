#!/bin/bash
# Builds tests/cusim/_build/libsseg_sim.so: the product's .cu sources compiled by g++ against the cusim model (no nvcc, no GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../semantic-segmentation-pytorch_b200/csrc"
# CUSIM_SANITIZE=address|thread: instrumented build in its own directory (run the tests with the matching runtime preloaded:
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0  /  LD_PRELOAD=$(gcc -print-file-name=libtsan.so))
# = out-of-bounds checking of every global / shared memory access of the kernels, resp. data-race detection between CUDA threads
SAN=""
OUT="$HERE/_build"
if [ -n "$CUSIM_SANITIZE" ]; then
  SAN="-fsanitize=$CUSIM_SANITIZE -fno-omit-frame-pointer"
  OUT="$HERE/_build/$CUSIM_SANITIZE"
fi
# CUSIM_COVERAGE=1: gcov-instrumented build (which lines of csrc/*.cu do the tests execute?) in _build/coverage
if [ -n "$CUSIM_COVERAGE" ]; then
  SAN="--coverage -O0"
  OUT="$HERE/_build/coverage"
fi
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
mkdir -p "$OUT"
exec 9>"$OUT/.lock"   # several test processes (xdist workers, torchrun ranks) may ask for the library at once: one builds
flock 9
FLAGS="$SAN -O2 -g -fno-strict-aliasing -std=c++17 -fPIC -pthread -w -I$HERE -I$CUDA_INC -include $HERE/cusim.h"
pids=()
for f in "$CSRC"/*.cu; do
  o="$OUT/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$CSRC" "$HERE" -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o")" ]; then
    g++ $FLAGS -x c++ -c "$f" -o "$o" &
    pids+=($!)
  fi
done
o="$OUT/selftest.o"
if [ ! -f "$o" ] || [ "$HERE/selftest.cu" -nt "$o" ] || [ "$HERE/cusim.h" -nt "$o" ]; then
  g++ $FLAGS -x c++ -c "$HERE/selftest.cu" -o "$o" &
  pids+=($!)
fi
o="$OUT/cusim_runtime.o"
if [ ! -f "$o" ] || [ "$HERE/cusim_runtime.cpp" -nt "$o" ] || [ "$HERE/cusim.h" -nt "$o" ] || [ "$HERE/cusim_ptx.h" -nt "$o" ]; then
  g++ $FLAGS -c "$HERE/cusim_runtime.cpp" -o "$o" &
  pids+=($!)
fi
for p in "${pids[@]}"; do wait "$p"; done
g++ $SAN -shared -pthread -Wl,-Bsymbolic -o "$OUT/libsseg_sim.so" "$OUT"/*.o -lrt
echo "$OUT/libsseg_sim.so"
