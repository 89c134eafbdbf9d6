python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_final.txt; cat gpurun_out/pytest_final.txt
timeout 100 tools/bin/api_bench_b200 b200 fhe > gpurun_out/api_bench_fhe3.txt 2>&1; grep '"logn": 17' gpurun_out/api_bench_fhe3.txt
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['reference_gpu_same_box']['value'], d['cpu_baseline']['value'])"
