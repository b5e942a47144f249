mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_box.txt
(timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15) > gpurun_out/r02a_pytest_gpu.log 2>&1
(timeout 600 python bench.py 2>gpurun_out/r02a_bench_default.err | tail -1) > gpurun_out/r02a_bench_default.json
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline 2>gpurun_out/r02a_bench_xbd.err | tail -1) > gpurun_out/r02a_bench_xbd1024.json
tail -5 gpurun_out/r02a_pytest_gpu.log; cut -c1-600 gpurun_out/r02a_bench_default.json; tail -3 gpurun_out/r02a_bench_default.err
