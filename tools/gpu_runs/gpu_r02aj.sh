mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_training.py -q -x 2>&1 | tail -n 45) > gpurun_out/r02aj_tests.log 2>&1
cat gpurun_out/r02aj_tests.log | cut -c1-250
