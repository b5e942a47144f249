mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "kernels or module" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -n 30) > gpurun_out/r02ad_tests.log 2>&1
cat gpurun_out/r02ad_tests.log
(timeout 300 python tools/bench_train_decoder.py 2>gpurun_out/r02ad.err | tail -n 1) > gpurun_out/r02_train_decoder_kernels.json
python -c "
import json; d=json.load(open('gpurun_out/r02_train_decoder_kernels.json'))
for r in d['kernels']: print(r['call'], round(r['fwd_ms'],4), round(r['bwd_ms'],4), round(r['fwd_tflops'],1), round(r['bwd_tflops'],1))
print(d['fwd_ms_total'], d['bwd_ms_total'])"
tail -n 3 gpurun_out/r02ad.err
