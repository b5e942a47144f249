mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -n 30) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -n 12 gpurun_out/r02_pytest_gpu.log | cut -c1-250
