mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "decoder or token" 2>&1 | tail -4) > gpurun_out/r02m_blocks.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -s -k "golden or define_G or tensor_core_modes or xbd_1024_golden or edge" 2>&1 | grep -E "parity.*tf32x3|passed|failed|rror" | tail -30) > gpurun_out/r02m_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --dump-kernels gpurun_out/r02m_kernels.json 2>gpurun_out/r02m_bench.err | tail -1) > gpurun_out/r02m_bench.json
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --dump-kernels gpurun_out/r02m_kernels_xbd.json 2>/dev/null | tail -1) > gpurun_out/r02m_bench_xbd.json
tail -n 3 gpurun_out/r02m_blocks.log; tail -n 22 gpurun_out/r02m_forward.log | cut -c1-200; cut -c1-150 gpurun_out/r02m_bench.json; cut -c1-150 gpurun_out/r02m_bench_xbd.json; tail -n 2 gpurun_out/r02m_bench.err
