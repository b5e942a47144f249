mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_ref_harness.py -q -x -s 2>&1 | grep -v Warning | grep -E "^\[train|passed|failed|Error|assert" | head -n 40 > gpurun_out/r02t_tests.log
cat gpurun_out/r02t_tests.log
(timeout 300 python tools/train_step.py --steps 20 2>gpurun_out/r02t.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native.json
(timeout 300 python tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02t.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graph --stock 2>>gpurun_out/r02t.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_graph.json
cut -c1-900 gpurun_out/r02_train_step_1gpu_native.json gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_graph.json
timeout 300 python tools/profile_train_step.py > gpurun_out/r02_train_step_profile_native.txt 2>>gpurun_out/r02t.err
tail -n 5 gpurun_out/r02t.err
