mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "split16" 2>&1 | tail -4) > gpurun_out/r02l_blocks.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -k "golden or define_G or full_size or xbd_1024_golden or graph" 2>&1 | tail -4) > gpurun_out/r02l_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02l_kernels.json 2>gpurun_out/r02l_bench.err | tail -1) > gpurun_out/r02l_bench.json
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02l_kernels_xbd.json 2>/dev/null | tail -1) > gpurun_out/r02l_bench_xbd.json
tail -n 3 gpurun_out/r02l_blocks.log; tail -n 3 gpurun_out/r02l_forward.log; cut -c1-150 gpurun_out/r02l_bench.json; cut -c1-150 gpurun_out/r02l_bench_xbd.json; tail -n 2 gpurun_out/r02l_bench.err
