mkdir -p gpurun_out
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 2>gpurun_out/r02am.err | tail -n 1) > gpurun_out/r02_bench_default_tf32x3.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_bench_default_tf32x3.json")); print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["parity"]["max_abs"]); print(d["training_step"])
EOF
tail -n 2 gpurun_out/r02am.err | cut -c1-200
