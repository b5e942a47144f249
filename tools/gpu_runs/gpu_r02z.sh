mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "training_step" 2>&1 | grep -E "^\[train|passed|failed|Error|assert" | head -n 40) > gpurun_out/r02z_tests.log 2>&1
cat gpurun_out/r02z_tests.log
