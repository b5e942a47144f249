mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
(timeout 400 $TR bench.py --gpus 2 --no-cpu-baseline --no-parity 2>gpurun_out/r02al.err | tail -n 1) > gpurun_out/r02_bench_levir256_2gpu.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_bench_levir256_2gpu.json")); print(d["value"], d["e2e"]["value"], d["strong_scaling"]["value"], d["strong_scaling"]["cuda_graph"]["value"]); print(d["training_step"])
EOF
tail -n 3 gpurun_out/r02al.err | cut -c1-300
