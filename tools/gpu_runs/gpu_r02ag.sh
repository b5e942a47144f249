mkdir -p gpurun_out
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02ag.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graph --foreach-adamw 2>>gpurun_out/r02ag.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph_foreach_adamw.json
grep -o '"step_ms": [0-9.]*\|loss_last": [0-9.]*' gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native_graph_foreach_adamw.json
tail -n 3 gpurun_out/r02ag.err | cut -c1-300
