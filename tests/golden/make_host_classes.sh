#!/bin/bash
# Regenerates tests/golden/host_classes.txt: tests/host_classes_probe.cu compiled against the reference's headers together with the
# reference's own CPU sources where they lie under /root/reference (build container only), n_inv_gpu masked (the reference never
# assigns it).  tests/test_host_classes.py::test_golden_is_what_the_reference_sources_print re-checks the committed file the same way.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=${REF:-/root/reference}
CUDA=${CUDA:-/usr/local/cuda}
T=$(mktemp -d)
g++ -O2 -std=c++17 -w -I$REF/src/include -I$CUDA/include -x c++ "$HERE/../host_classes_probe.cu" $REF/src/lib/common/common.cu \
    $REF/src/lib/common/nttparameters.cu $REF/src/lib/ntt_merge/ntt_cpu.cu $REF/src/lib/ntt_4step/ntt_4step_cpu.cu -o $T/probe \
    -L$CUDA/lib64 -lcudart_static -ldl -lrt -lpthread
$T/probe | sed -E 's/n_inv_gpu=[0-9]+/n_inv_gpu=*/' > "$HERE/host_classes.txt"
rm -rf $T
wc -l "$HERE/host_classes.txt"
