mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -s 2>&1 | grep -v Warning | grep -E "^\[train|passed|failed|Error|assert" | head -n 40 > gpurun_out/r02u_tests.log
cat gpurun_out/r02u_tests.log
(timeout 300 python tools/bench_train_decoder.py 2>gpurun_out/r02u.err | tail -n 1) > gpurun_out/r02_train_decoder_kernels.json
python -c "
import json; d=json.load(open('gpurun_out/r02_train_decoder_kernels.json'))
for r in d['kernels']: print(r['call'], round(r['fwd_ms'],4), round(r['bwd_ms'],4))
print(d['fwd_ms_total'], d['bwd_ms_total'])"
(timeout 300 python tools/train_step.py --steps 20 --graph 2>>gpurun_out/r02u.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph.json
(timeout 300 python tools/train_step.py --steps 20 --graph --channels-last 2>>gpurun_out/r02u.err | grep '^{' | tail -n 1) > gpurun_out/r02_train_step_1gpu_native_graph_cl.json
cut -c1-500 gpurun_out/r02_train_step_1gpu_native_graph.json gpurun_out/r02_train_step_1gpu_native_graph_cl.json
timeout 300 python tools/profile_train_step.py --channels-last > gpurun_out/r02_train_step_profile_native_cl.txt 2>>gpurun_out/r02u.err
tail -n 5 gpurun_out/r02u.err
