mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "split16" 2>&1 | tail -30) > gpurun_out/r02d_blocks.log 2>&1
(timeout 1200 python -m pytest tests/test_gpu_forward.py -q -x -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | tail -60) > gpurun_out/r02d_forward.log 2>&1
for cg in 1 0 2; do
(DAHITRA_TC3_CG=$cg timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02d_kernels_cg$cg.json 2>gpurun_out/r02d_bench_cg$cg.err | tail -1) > gpurun_out/r02d_bench_cg$cg.json
done
tail -8 gpurun_out/r02d_blocks.log; tail -6 gpurun_out/r02d_forward.log; for cg in 1 0 2; do cut -c1-160 gpurun_out/r02d_bench_cg$cg.json; done; tail -3 gpurun_out/r02d_bench_cg1.err
