N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
(timeout 600 $TR tools/train_step.py --steps 20 2>gpurun_out/r02m3_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu.json
(timeout 600 $TR tools/train_step.py --steps 20 --freeze-unused 2>>gpurun_out/r02m3_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_frozen_unused.json
(timeout 600 $TR tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02m3_${N}gpu.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_graph.json
python - <<EOF
import json
for f in ("gpurun_out/r02_train_step_${N}gpu.json","gpurun_out/r02_train_step_${N}gpu_frozen_unused.json","gpurun_out/r02_train_step_${N}gpu_graph.json"):
    d=json.load(open(f)); print({k:d.get(k) for k in ("ddp","step_ms","step_ms_without_allreduce","pairs_per_s","exposed_allreduce_share","allreduce_alone_ms","replicas_equal","loss_last","params_without_grad")})
EOF
tail -n 3 gpurun_out/r02m3_${N}gpu.err
