#!/bin/bash
# tools/build_variant.sh NAME "-DMACRO=.. ..." : builds gpu_ntt_b200/lib/libgpuntt_b200_NAME.so for A/B timing (GPUNTT_B200_LIB=...)
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -cudart static \
    -I"$ROOT/include" -I"$ROOT/gpu_ntt_b200/csrc" $2 -o "$ROOT/gpu_ntt_b200/lib/libgpuntt_b200_$1.so" "$ROOT"/gpu_ntt_b200/csrc/*.cu
echo "built variant $1"
