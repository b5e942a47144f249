mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "graphed" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -n 30) > gpurun_out/r02af_tests.log 2>&1
cat gpurun_out/r02af_tests.log
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02af.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graphed-route 2>>gpurun_out/r02af.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graphed_route.json
grep -o '"step_ms": [0-9.]*' gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native_graphed_route.json
tail -n 3 gpurun_out/r02af.err | cut -c1-300
