#!/bin/bash
# one multi-GPU gpurun call: configs 2, 3, 5 (and 4 at N = 8) under torchrun on N GPUs of one box
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
(timeout 600 $TR bench.py --gpus $N --no-cpu-baseline --no-parity 2>gpurun_out/r02_bench_${N}gpu.err | tail -n 1) > gpurun_out/r02_bench_levir256_${N}gpu.json
(timeout 600 $TR bench.py --gpus $N --workload xbd1024 --no-cpu-baseline --no-parity 2>>gpurun_out/r02_bench_${N}gpu.err | tail -n 1) > gpurun_out/r02_bench_xbd1024_${N}gpu.json
(timeout 600 $TR tools/bench_tiled1024.py --batches 8,64 2>>gpurun_out/r02_bench_${N}gpu.err | grep '^{') > gpurun_out/r02_tiled1024_bf16_${N}gpu.jsonl
(timeout 600 $TR tools/train_step.py --steps 20 2>>gpurun_out/r02_bench_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu.json
(timeout 600 $TR tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02_bench_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_graph.json
if [ "$N" -ge 2 ]; then (timeout 600 python -m pytest tests/test_gpu_forward.py -q -x -k "non_current_device" 2>&1 | tail -n 2) > gpurun_out/r02_non_current_device_${N}gpu.log; fi
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv,noheader > gpurun_out/r02_box_${N}gpu.txt
for f in r02_bench_levir256_${N}gpu.json r02_bench_xbd1024_${N}gpu.json; do cut -c1-220 gpurun_out/$f; done
cat gpurun_out/r02_tiled1024_bf16_${N}gpu.jsonl | cut -c1-300; cut -c1-500 gpurun_out/r02_train_step_${N}gpu.json gpurun_out/r02_train_step_${N}gpu_graph.json; cat gpurun_out/r02_non_current_device_${N}gpu.log 2>/dev/null; tail -n 3 gpurun_out/r02_bench_${N}gpu.err
