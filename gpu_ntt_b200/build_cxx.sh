#!/bin/bash
# Builds the C++ link surface without CMake: gpu_ntt_b200/lib/libntt-1.0.a (what the CMake target `ntt`
# produces) and, where the reference tree is present, the reference's OWN example programs compiled unchanged
# against include/gpuntt/ + that archive (tests/_dropin/, git-ignored; the binaries travel to the GPU box).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$ROOT/include -I$HERE/csrc"
OBJ="$HERE/lib/obj"
mkdir -p "$OBJ"
# the kernels and the C ABI (csrc/*.cu) are the objects build.sh compiles for the shared library (same flags, -fPIC): reuse them
bash "$HERE/build.sh" > /dev/null
rm -f "$OBJ"/merge_*.o
for f in cxx/common cxx/nttparameters cxx/ntt_cpu cxx/ntt_api cxx/ntt_4step_api; do
    o="$OBJ/$(basename $f).o"
    if [ ! -f "$o" ] || [ "$HERE/$f.cu" -nt "$o" ] || [ -n "$(find "$HERE/csrc" "$ROOT/include" -newer "$o" -name '*.*h' -o -newer "$o" -name '*.inl' | head -1)" ]; then
        $NVCC $FLAGS -c -o "$o" "$HERE/$f.cu" &
    fi
done
wait
rm -f "$HERE/lib/libntt-1.0.a"
ar rcs "$HERE/lib/libntt-1.0.a" "$HERE/lib/obj_so"/merge_*.o "$OBJ"/*.o
echo "built $HERE/lib/libntt-1.0.a"
for ex in gpu_merge_examples gpu_4step_examples; do
    if [ -f "$ROOT/examples/$ex.cu" ]; then
        $NVCC $FLAGS -o "$HERE/lib/$ex" "$ROOT/examples/$ex.cu" "$HERE/lib/libntt-1.0.a" -cudart static &
    fi
done
REF=${REF:-/root/reference}
if [ -d "$REF/example" ]; then
    mkdir -p "$ROOT/tests/_dropin"
    for pair in gpu_merge_ntt_examples:ntt_merge/test_merge_ntt.cu gpu_merge_intt_examples:ntt_merge/test_merge_intt.cu \
                cpu_merge_ntt_examples:ntt_merge/test_cpu_merge_ntt.cu gpu_4step_ntt_examples:ntt_4step/test_4step_ntt.cu \
                gpu_4step_intt_examples:ntt_4step/test_4step_intt.cu cpu_4step_ntt_examples:ntt_4step/test_cpu_4step_ntt.cu; do
        exe=${pair%%:*}; src=${pair##*:}
        $NVCC $FLAGS -w -o "$ROOT/tests/_dropin/$exe" "$REF/example/$src" "$HERE/lib/libntt-1.0.a" -cudart static &
    done
fi
wait
echo "examples built"
