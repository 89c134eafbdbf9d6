#!/bin/bash
# Builds tools/api_bench.cu twice: against the reference's headers + its own GPU sources compiled for sm_100
# (oracle/_ref/libntt_ref_gpu.a, `make -C oracle refgpu`; only where /root/reference exists) and against this
# repository's include/gpuntt + libntt-1.0.a.  Binaries go to tools/bin/ (git-ignored; they travel to the GPU box).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
REF=${REF:-/root/reference}
mkdir -p "$HERE/bin"
[ -f "$ROOT/gpu_ntt_b200/lib/libntt-1.0.a" ] || bash "$ROOT/gpu_ntt_b200/build_cxx.sh"
$NVCC -O3 -std=c++17 -w -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -o "$HERE/bin/api_bench_b200" \
    "$HERE/api_bench.cu" "$ROOT/gpu_ntt_b200/lib/libntt-1.0.a" -cudart static &
if [ -d "$REF/src" ]; then
    make -s -C "$ROOT/oracle" refgpu
    $NVCC -O3 -std=c++17 -w -gencode arch=compute_100,code=sm_100 -I"$REF/src/include" -o "$HERE/bin/api_bench_reference" \
        "$HERE/api_bench.cu" "$ROOT/oracle/_ref/libntt_ref_gpu.a" -cudart static &
fi
wait
ls -la "$HERE/bin"/api_bench_*
