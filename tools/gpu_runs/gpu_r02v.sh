mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 8) > gpurun_out/r02v_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/r02v_pytest_gpu.log
(timeout 300 python tools/train_step.py --steps 20 --graph 2>gpurun_out/r02v.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 2>>gpurun_out/r02v.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native.json
(timeout 300 python tools/train_step.py --steps 20 --graph --nchw 2>>gpurun_out/r02v.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph_nchw.json
cut -c1-420 gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native.json gpurun_out/r02_train_step_1gpu_native_graph_nchw.json
(timeout 600 python bench.py --no-cpu-baseline 2>>gpurun_out/r02v.err | tail -n 1) > gpurun_out/r02v_bench_default.json
cut -c1-300 gpurun_out/r02v_bench_default.json
tail -n 5 gpurun_out/r02v.err
