mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -s -k "split16 or stem or decoder" 2>&1 | tail -60) > gpurun_out/r02c_blocks.log 2>&1
(timeout 1200 python -m pytest tests/test_gpu_forward.py -q -x -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | tail -60) > gpurun_out/r02c_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/r02c_kernels_tf32x3.json 2>gpurun_out/r02c_bench.err | tail -1) > gpurun_out/r02c_bench_default.json
(DAHITRA_TC3_STREAM=1 timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02c_kernels_stream.json 2>/dev/null | tail -1) > gpurun_out/r02c_bench_stream.json
tail -8 gpurun_out/r02c_blocks.log; tail -12 gpurun_out/r02c_forward.log; cut -c1-200 gpurun_out/r02c_bench_default.json; cut -c1-200 gpurun_out/r02c_bench_stream.json; tail -3 gpurun_out/r02c_bench.err
