mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_forward.py -q -x 2>&1 | tail -n 3) > gpurun_out/r02o_forward.log 2>&1
rm -f gpurun_out/r02o_latency.jsonl
for b in 1 4 8; do for x in 0 65536; do (timeout 300 python tools/latency_small_batch.py --batch $b --extra-flags $x 2>/dev/null | tail -n 1) >> gpurun_out/r02o_latency.jsonl; done; done
tail -n 3 gpurun_out/r02o_forward.log; cut -c1-230 gpurun_out/r02o_latency.jsonl
