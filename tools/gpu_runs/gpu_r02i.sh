mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "decoder" 2>&1 | tail -4) > gpurun_out/r02i_blocks.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -s -k "golden or define_G or tensor_core_modes or xbd_1024_golden" 2>&1 | grep -E "parity|passed|failed|rror" | tail -30) > gpurun_out/r02i_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/r02i_kernels.json 2>gpurun_out/r02i_bench.err | tail -1) > gpurun_out/r02i_bench.json
tail -4 gpurun_out/r02i_blocks.log; tail -12 gpurun_out/r02i_forward.log; cut -c1-150 gpurun_out/r02i_bench.json; tail -2 gpurun_out/r02i_bench.err
