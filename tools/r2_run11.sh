#!/bin/bash
# round 2, GPU pass 11: inverse last round with n^-1 folded into the twiddles + twiddle-1 butterflies; 4-step forms; full suite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -8 gpurun_out/pytest_gpu.txt
timeout 900 python tools/perf_configs.py > gpurun_out/perf_configs.jsonl 2> gpurun_out/perf_err.txt; tail -3 gpurun_out/perf_err.txt; cut -c1-250 gpurun_out/perf_configs.jsonl | head -14
sed -n '/^cat > \/tmp\/san_4step.py/,/^PY$/p' tools/r2_run10.sh > /tmp/mk_san.sh; bash /tmp/mk_san.sh
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/san_4step.py > gpurun_out/sanitizer_4step_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Traceback|assert" gpurun_out/sanitizer_4step_$tool.txt | head -10; grep -c "^ok " gpurun_out/sanitizer_4step_$tool.txt
done
sed -n '/^cat > \/tmp\/perf4.py/,/^PY$/p' tools/r2_run10.sh > /tmp/mk_perf4.sh; bash /tmp/mk_perf4.sh
timeout 600 python /tmp/perf4.py > gpurun_out/perf_4step.jsonl 2> gpurun_out/perf_4step_err.txt; tail -3 gpurun_out/perf_4step_err.txt; cut -c1-330 gpurun_out/perf_4step.jsonl | head -8
