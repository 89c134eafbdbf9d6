#!/bin/bash
# Builds the C-ABI shared library IN-TREE for sm_100a (no GPU needed: nvcc cross-compiles).
# Translation units are compiled in parallel into lib/obj_so/ and only when stale.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${GPUNTT_SO_NAME:-libgpuntt_b200.so}
OBJ="$HERE/lib/obj_so${GPUNTT_OBJ_TAG}"
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$ROOT/include -I$HERE/csrc ${GPUNTT_NVCC_EXTRA}"
pids=()
for src in "$HERE"/csrc/*.cu; do
    o="$OBJ/$(basename "${src%.cu}").o"
    if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ -n "$(find "$HERE/csrc" "$ROOT/include" \( -name '*.cuh' -o -name '*.h' -o -name '*.inl' \) -newer "$o" | head -1)" ] || [ -n "${GPUNTT_FORCE}" ]; then
        $NVCC $FLAGS -c -o "$o" "$src" &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$HERE/lib/$OUT" "$OBJ"/*.o
echo "built $HERE/lib/$OUT"
