mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "classifier" 2>&1 | tail -5) > gpurun_out/r02h_blocks.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -k "golden or five_class or full_size or xbd_1024_golden" 2>&1 | tail -5) > gpurun_out/r02h_forward.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02h_kernels.json 2>gpurun_out/r02h_bench.err | tail -1) > gpurun_out/r02h_bench.json
(DAHITRA_CLS_V1=1 timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02h_kernels_clsv1.json 2>/dev/null | tail -1) > gpurun_out/r02h_bench_clsv1.json
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02h_kernels_xbd.json 2>/dev/null | tail -1) > gpurun_out/r02h_bench_xbd.json
(DAHITRA_CLS_V1=1 timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02h_kernels_xbd_clsv1.json 2>/dev/null | tail -1) > gpurun_out/r02h_bench_xbd_clsv1.json
tail -3 gpurun_out/r02h_blocks.log gpurun_out/r02h_forward.log; for f in r02h_bench r02h_bench_clsv1 r02h_bench_xbd r02h_bench_xbd_clsv1; do cut -c1-150 gpurun_out/$f.json; done; tail -2 gpurun_out/r02h_bench.err
