#!/bin/bash
# tools/build_variant.sh NAME "-DMACRO=.. ..." : builds gpu_ntt_b200/lib/libgpuntt_b200_NAME.so for A/B timing (GPUNTT_B200_LIB=...)
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
GPUNTT_SO_NAME="libgpuntt_b200_$1.so" GPUNTT_OBJ_TAG="_$1" GPUNTT_NVCC_EXTRA="$2" bash "$ROOT/gpu_ntt_b200/build.sh"
echo "built variant $1"
