mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "graph" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -n 30) > gpurun_out/r02aa_tests.log 2>&1
cat gpurun_out/r02aa_tests.log
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02aa.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
cut -c1-900 gpurun_out/r02_train_step_1gpu_native_graph.json; tail -n 5 gpurun_out/r02aa.err
