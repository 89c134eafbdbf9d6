#!/bin/bash
# round 2, GPU pass 12: A/B of the inverse last-round change (base = the commit before it) + sanitizer over the new 4-step forms
mkdir -p gpurun_out
timeout 900 python tools/ab_cases.py --rounds 2 --cases c2inv,c2neginv,c3inv,c4inv,c4invref,small,big gpu_ntt_b200/lib/libgpuntt_b200_base.so gpu_ntt_b200/lib/libgpuntt_b200.so > gpurun_out/ab_inverse.jsonl 2> gpurun_out/ab_err.txt; tail -3 gpurun_out/ab_err.txt; cat gpurun_out/ab_inverse.jsonl
sed -n '/^cat > \/tmp\/san_4step.py/,/^PY$/p' tools/r2_run10.sh > /tmp/mk_san.sh; bash /tmp/mk_san.sh
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python /tmp/san_4step.py > gpurun_out/sanitizer_4step_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Traceback|assert" gpurun_out/sanitizer_4step_$tool.txt | head -10; grep -c "^ok " gpurun_out/sanitizer_4step_$tool.txt
done
