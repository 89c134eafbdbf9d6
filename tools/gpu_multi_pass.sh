#!/bin/bash
# round 2, multi-GPU pass (8 GPUs of one box): host-link probe at 1/2/4/8 concurrent GPUs, two-GPU tests, bench at N = 8 and 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/multi_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/multi_gpus.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" >> gpurun_out/multi_gpus.txt 2>&1
timeout 600 python tools/pcie_probe_multi.py > gpurun_out/pcie_probe_multi.txt 2> gpurun_out/pcie_probe_err.txt; cat gpurun_out/pcie_probe_multi.txt; tail -2 gpurun_out/pcie_probe_err.txt
timeout 600 python -m pytest tests/test_multi_device_gpu.py -q > gpurun_out/pytest_multi.txt 2>&1; tail -3 gpurun_out/pytest_multi.txt
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu_err.txt; cut -c1-600 gpurun_out/bench_${n}gpu.json; tail -2 gpurun_out/bench_${n}gpu_err.txt
done
