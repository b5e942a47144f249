mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py tests/test_ref_harness.py -q -s -k "graphed or harness or trainer" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -n 30) > gpurun_out/r02ae_tests.log 2>&1
cat gpurun_out/r02ae_tests.log
(timeout 300 python tools/train_step.py --steps 20 --graphed-route 2>gpurun_out/r02ae.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graphed_route.json
(timeout 300 python tools/train_step.py --steps 20 2>>gpurun_out/r02ae.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native.json
cut -c1-700 gpurun_out/r02_train_step_1gpu_native_graphed_route.json; grep -o '"step_ms": [0-9.]*' gpurun_out/r02_train_step_1gpu_native.json
tail -n 5 gpurun_out/r02ae.err
