mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -s 2>&1 | grep -v Warning | grep -E "^\[train|passed|failed|Error|assert" | head -n 40 > gpurun_out/r02s_tests.log
cat gpurun_out/r02s_tests.log
(timeout 300 python tools/train_ablate.py 2>gpurun_out/r02s.err | tail -n 1) > gpurun_out/r02_train_step_ablation.json
cat gpurun_out/r02_train_step_ablation.json; tail -n 5 gpurun_out/r02s.err
(timeout 300 python tools/bench_train_decoder.py 2>>gpurun_out/r02s.err | tail -n 1) > gpurun_out/r02_train_decoder_kernels.json
cat gpurun_out/r02_train_decoder_kernels.json
