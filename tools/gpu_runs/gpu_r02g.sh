mkdir -p gpurun_out
F=$(python -c "from dahitra_b200.engine import MODES; print(MODES['tf32x3'] | 32768)")
(DAHITRA_FLAGS=$F timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -k "tf32x3 and (golden or full_size or graph or streams or define_G or xbd_1024_golden)" 2>&1 | tail -5) > gpurun_out/r02g_forward_pdl.log 2>&1
for b in 4 8; do for x in 0 32768 33024 256; do
(timeout 300 python tools/latency_small_batch.py --batch $b --extra-flags $x 2>/dev/null | tail -1) >> gpurun_out/r02g_latency.jsonl
done; done
(timeout 600 python bench.py --no-cpu-baseline --no-parity --flags $F --dump-kernels gpurun_out/r02g_kernels_pdl.json 2>gpurun_out/r02g_bench_pdl.err | tail -1) > gpurun_out/r02g_bench_pdl.json
(timeout 600 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | tail -1) > gpurun_out/r02g_bench_nopdl.json
tail -5 gpurun_out/r02g_forward_pdl.log; cat gpurun_out/r02g_latency.jsonl; cut -c1-160 gpurun_out/r02g_bench_pdl.json; cut -c1-160 gpurun_out/r02g_bench_nopdl.json; tail -2 gpurun_out/r02g_bench_pdl.err
