#!/bin/bash
# Builds the C-ABI shared library IN-TREE for sm_100a (no GPU needed: nvcc cross-compiles).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
mkdir -p "$HERE/lib"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC -shared -cudart static --threads 0 -I"$ROOT/include" -I"$HERE/csrc" \
    ${GPUNTT_NVCC_EXTRA} \
    -o "$HERE/lib/libgpuntt_b200.so" "$HERE"/csrc/*.cu
echo "built $HERE/lib/libgpuntt_b200.so"
