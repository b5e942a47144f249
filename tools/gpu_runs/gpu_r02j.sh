mkdir -p gpurun_out
(timeout 600 python tools/train_step.py --steps 20 2>gpurun_out/r02j_train.err | grep '^{' | tail -1) > gpurun_out/r02_train_step_1gpu.json
(timeout 600 python tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02j_train.err | grep '^{' | tail -1) > gpurun_out/r02_train_step_1gpu_graph.json
cat gpurun_out/r02_train_step_1gpu.json gpurun_out/r02_train_step_1gpu_graph.json; tail -5 gpurun_out/r02j_train.err
