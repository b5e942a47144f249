mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_forward.py tests/test_metrics.py -q -x -k "classifier or xbd or metrics or confusion" 2>&1 | tail -n 4) > gpurun_out/r02ac_tests.log 2>&1
cat gpurun_out/r02ac_tests.log
(timeout 600 python bench.py --workload xbd1024 --no-cpu-baseline --dump-kernels gpurun_out/r02ac_kernels_xbd.json 2>gpurun_out/r02ac.err | tail -n 1) > gpurun_out/r02ac_bench_xbd.json
cut -c1-250 gpurun_out/r02ac_bench_xbd.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02ac_kernels_xbd.json"))
ks=d if isinstance(d,list) else d.get("kernels",d)
for k in ks:
    if isinstance(k,dict) and k.get("name") in ("classifier","conv_layer2","stem","conv_layer2_0.0"): print(k)
EOF
tail -n 3 gpurun_out/r02ac.err
