mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -n 6) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/r02_pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2) > gpurun_out/r02_smoke.log 2>&1
cat gpurun_out/r02_smoke.log
