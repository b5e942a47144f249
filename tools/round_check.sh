#!/bin/bash
# Round-end style check in one gpurun call (release build): full GPU suite, smoke(), default bench (+ kernel dump), xBD bench,
# reference arm, small-batch latency, precision ablation.  Everything lands in gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${TAG}_smoke.log 2>&1
(timeout 600 python bench.py --dump-kernels gpurun_out/${TAG}_kernels_tf32x3.json 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_default_tf32x3.json
(timeout 600 python bench.py --workload xbd1024 --dump-kernels gpurun_out/${TAG}_kernels_xbd1024_tf32x3.json 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_xbd1024_tf32x3.json
for m in tf32x3_fp32act f16 bf16 tf32; do
(timeout 600 python bench.py --mode $m --no-cpu-baseline --no-parity 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_$m.json
done
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_reference_cpu.json
rm -f gpurun_out/${TAG}_latency_small_batch.jsonl
for b in 1 4 8 16; do (timeout 300 python tools/latency_small_batch.py --batch $b 2>/dev/null | tail -1) >> gpurun_out/${TAG}_latency_small_batch.jsonl; done
(timeout 600 python tools/ablate_flags.py 2>/dev/null) > gpurun_out/${TAG}_precision_ablation.txt
# training step (configs[3]) at one GPU: eager, graphed (dahitra_b200.train_graph), the kernels alone, the ablation
(timeout 300 python tools/train_step.py --steps 20 2>/dev/null | grep '^{' | tail -n 1) > gpurun_out/${TAG}_train_step_1gpu_native.json
(timeout 300 python tools/train_step.py --steps 20 --graph 2>/dev/null | grep '^{' | tail -n 1) > gpurun_out/${TAG}_train_step_1gpu_native_graph.json
(timeout 300 python tools/bench_train_decoder.py 2>/dev/null | tail -n 1) > gpurun_out/${TAG}_train_decoder_kernels.json
(timeout 300 python tools/train_ablate.py 2>/dev/null | tail -n 1) > gpurun_out/${TAG}_train_step_ablation.json
tail -3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_smoke.log
for f in default_tf32x3 xbd1024_tf32x3 tf32x3_fp32act f16 bf16 tf32 reference_cpu; do cut -c1-200 gpurun_out/${TAG}_bench_$f.json; done
cat gpurun_out/${TAG}_latency_small_batch.jsonl | cut -c1-260
for f in train_step_1gpu_native train_step_1gpu_native_graph train_step_ablation; do cut -c1-600 gpurun_out/${TAG}_$f.json | grep -o '"step_ms": [0-9.]*\|full_[a-z_]*": [0-9.]*' ; done
