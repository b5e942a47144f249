N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
(timeout 300 $TR tools/train_step.py --steps 30 --graph 2>gpurun_out/r02m4_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_graph.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_train_step_${N}gpu_graph.json")); print({k:d.get(k) for k in ("step_ms","step_ms_without_allreduce","pairs_per_s","pairs_per_s_per_gpu","exposed_allreduce_share","allreduce_alone_ms","replicas_equal","loss_first","loss_last","optimizer")})
EOF
tail -n 3 gpurun_out/r02m4_${N}gpu.err | cut -c1-300
