mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -s -k "split16" 2>&1 | tail -60) > gpurun_out/r02b_blocks.log 2>&1
(timeout 1200 python -m pytest tests/test_gpu_forward.py -q -x -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | tail -60) > gpurun_out/r02b_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/r02b_kernels_tf32x3.json 2>gpurun_out/r02b_bench.err | tail -1) > gpurun_out/r02b_bench_default.json
(timeout 600 python bench.py --no-cpu-baseline --no-parity --mode tf32x3_fp32act 2>/dev/null | tail -1) > gpurun_out/r02b_bench_fp32act.json
tail -25 gpurun_out/r02b_blocks.log; tail -30 gpurun_out/r02b_forward.log; cut -c1-300 gpurun_out/r02b_bench_default.json; tail -3 gpurun_out/r02b_bench.err
