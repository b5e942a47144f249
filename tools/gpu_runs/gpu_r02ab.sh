mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -s -k "graph" 2>&1 | grep -E "^\[|passed|failed|Error|assert" | head -n 30) > gpurun_out/r02ab_tests.log 2>&1
cat gpurun_out/r02ab_tests.log
for i in 1 2; do
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02ab.err | grep '^{' | tail -n 1) | grep -o '"step_ms": [0-9.]*'
(timeout 300 python tools/_train_step_prev.py --steps 20 --graph 2>>gpurun_out/r02ab.err | grep '^{' | tail -n 1) | grep -o '"step_ms": [0-9.]*'
done
tail -n 5 gpurun_out/r02ab.err
