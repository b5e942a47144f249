#!/bin/bash
# Round-end style check in one gpurun call: full GPU suite, smoke(), default bench, tf32 bench, reference arm,
# ncu launch list + full capture of the dominant kernels (default mode).  Everything lands in gpurun_out/.
TAG=${1:-r01}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4) > gpurun_out/pytest_gpu_$TAG.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/smoke_$TAG.log 2>&1
(timeout 600 python bench.py 2>&1 | tail -1) > gpurun_out/bench_default_$TAG.json 2>&1
(timeout 600 python bench.py --mode tf32 --no-cpu-baseline --dump-kernels gpurun_out/kernels_tf32_$TAG.json 2>&1 | tail -1) > gpurun_out/bench_tf32_$TAG.json 2>&1
(timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/kernels_tf32x3_$TAG.json 2>&1 | tail -1) > /dev/null 2>&1
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_xbd1024_$TAG.json 2>&1
(timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/bench_reference_$TAG.json 2>&1
(timeout 300 python tools/latency_small_batch.py 2>&1 | tail -1) > gpurun_out/latency_b4_$TAG.json 2>&1
(timeout 300 python tools/latency_small_batch.py --mode tf32 2>&1 | tail -1) >> gpurun_out/latency_b4_$TAG.json 2>&1
tools/ncu_capture.sh tf32x3 $TAG "conv_tc2_kernel pixel_decoder_tc_kernel stem_f16_kernel" > gpurun_out/ncu_capture_stdout_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log gpurun_out/smoke_$TAG.log
for f in default tf32 xbd1024 reference; do cut -c1-330 gpurun_out/bench_${f}_$TAG.json; done
cat gpurun_out/latency_b4_$TAG.json
