N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
mkdir -p gpurun_out
(timeout 600 $TR tools/train_step.py --steps 20 2>gpurun_out/r02q_train_${N}.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu.json
(timeout 600 $TR tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02q_train_${N}.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_${N}gpu_graph.json
cut -c1-700 gpurun_out/r02_train_step_${N}gpu.json; cut -c1-900 gpurun_out/r02_train_step_${N}gpu_graph.json; tail -n 5 gpurun_out/r02q_train_${N}.err
