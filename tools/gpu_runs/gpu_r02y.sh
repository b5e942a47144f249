mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "training_step" 2>&1 | grep -E "^\[train|passed|failed|Error|assert" | head -n 40) > gpurun_out/r02y_tests.log 2>&1
cat gpurun_out/r02y_tests.log
for i in 1 2; do
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02y.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graph --two-pass-trunk 2>>gpurun_out/r02y.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph_two_pass_trunk.json
cat gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native_graph_two_pass_trunk.json | grep -o '"trunk.*step_ms": [0-9.]*'
done
(timeout 300 python tools/train_step.py --steps 20 2>>gpurun_out/r02y.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native.json
grep -o '"step_ms": [0-9.]*' gpurun_out/r02_train_step_1gpu_native.json
tail -n 3 gpurun_out/r02y.err
