mkdir -p gpurun_out
(timeout 1800 python -m pytest tests/test_gpu_training.py tests/test_gpu_forward.py -q -x -s -k "train or scheduling or c_abi or graph" 2>&1 | grep -E "^\[train|passed|failed|Error|assert" | head -n 40) > gpurun_out/r02w_tests.log 2>&1
cat gpurun_out/r02w_tests.log
for i in 1 2; do
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02w.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graph --nchw 2>>gpurun_out/r02w.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph_nchw.json
cut -c1-420 gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native_graph_nchw.json | grep -o '"memory_format.*step_ms": [0-9.]*'
done
(timeout 300 python tools/train_step.py --steps 20 2>>gpurun_out/r02w.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native.json
grep -o '"step_ms": [0-9.]*' gpurun_out/r02_train_step_1gpu_native.json
timeout 300 python tools/profile_train_step.py > gpurun_out/r02_train_step_profile_native.txt 2>>gpurun_out/r02w.err
tail -n 3 gpurun_out/r02w.err
