# reduced multi-GPU check: configs[1] (weak + strong scaling, eager and CUDA graph) and configs[3] (training step: DDP eager, flat-buffer graph)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
(timeout 600 $TR bench.py --gpus $N --no-cpu-baseline --no-parity 2>gpurun_out/r02_bench_${N}gpu.err | tail -n 1) > gpurun_out/r02_bench_levir256_${N}gpu.json
(timeout 600 $TR tools/train_step.py --steps 20 2>>gpurun_out/r02_bench_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu.json
(timeout 600 $TR tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02_bench_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_graph.json
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv,noheader > gpurun_out/r02_box_${N}gpu.txt
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_bench_levir256_${N}gpu.json")); print(d["value"], d["e2e"]["value"], d["strong_scaling"])
for f in ("gpurun_out/r02_train_step_${N}gpu.json","gpurun_out/r02_train_step_${N}gpu_graph.json"):
    d=json.load(open(f)); print({k:d[k] for k in ("step_ms","pairs_per_s","pairs_per_s_per_gpu","exposed_allreduce_share","allreduce_alone_ms","replicas_equal","loss_first","loss_last")})
EOF
tail -n 3 gpurun_out/r02_bench_${N}gpu.err
