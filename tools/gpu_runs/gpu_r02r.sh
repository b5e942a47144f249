mkdir -p gpurun_out
(timeout 300 python tools/train_ablate.py 2>gpurun_out/r02r.err | tail -n 1) > gpurun_out/r02_train_step_ablation.json
cat gpurun_out/r02_train_step_ablation.json; tail -n 5 gpurun_out/r02r.err
