mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py -q -x -k "split16 or token or stem" 2>&1 | tail -n 3) > gpurun_out/r02n_blocks.log 2>&1
(timeout 1200 python -m pytest tests/test_gpu_forward.py -q -x 2>&1 | tail -n 3) > gpurun_out/r02n_forward.log 2>&1
rm -f gpurun_out/r02n_latency.jsonl
for b in 1 4 8 16; do (timeout 300 python tools/latency_small_batch.py --batch $b 2>/dev/null | tail -n 1) >> gpurun_out/r02n_latency.jsonl; done
(timeout 600 python bench.py --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02n_kernels.json 2>gpurun_out/r02n_bench.err | tail -n 1) > gpurun_out/r02n_bench.json
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --no-parity --dump-kernels gpurun_out/r02n_kernels_xbd.json 2>/dev/null | tail -n 1) > gpurun_out/r02n_bench_xbd.json
tail -n 3 gpurun_out/r02n_blocks.log; tail -n 3 gpurun_out/r02n_forward.log; cut -c1-230 gpurun_out/r02n_latency.jsonl; cut -c1-150 gpurun_out/r02n_bench.json; cut -c1-150 gpurun_out/r02n_bench_xbd.json; tail -n 2 gpurun_out/r02n_bench.err
