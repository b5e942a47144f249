mkdir -p gpurun_out
(timeout 600 python tools/train_step.py --steps 20 2>gpurun_out/r02p_train.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu.json
(timeout 600 python tools/train_step.py --steps 20 --as-written 2>>gpurun_out/r02p_train.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_as_written.json
(timeout 600 python tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02p_train.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_graph.json
(timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_ref_harness.py -q -x -k "autograd or training or harness" 2>&1 | tail -n 3) > gpurun_out/r02p_tests.log
cut -c1-420 gpurun_out/r02_train_step_1gpu.json gpurun_out/r02_train_step_1gpu_as_written.json gpurun_out/r02_train_step_1gpu_graph.json; cat gpurun_out/r02p_tests.log; tail -n 3 gpurun_out/r02p_train.err
