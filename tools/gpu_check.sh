#!/bin/bash
# One gpurun call: parity tests, tcgen05 bring-up tests (isolated process, bounded), bench in both modes,
# ncu launch list and one full capture of the dominant kernel.  Everything lands in gpurun_out/.
#   usage: tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "not tcgen05" 2>&1 | tail -15) > gpurun_out/blocks_$TAG.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -k "not tensor_core" 2>&1 | grep -E "parity|passed|failed|FAILED|Error" | tail -40) > gpurun_out/forward_$TAG.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_blocks.py -q -m gpu -s -k "tcgen05" 2>&1 | grep -E "\[tc\]|passed|failed|FAILED|Error|error|max\|d" | tail -60) > gpurun_out/tc_blocks_$TAG.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_forward.py -q -m gpu -s -k "tensor_core" 2>&1 | grep -E "parity|passed|failed|FAILED|Error|error" | tail -20) > gpurun_out/tc_forward_$TAG.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1) > gpurun_out/bench_f0_$TAG.json 2>&1
(timeout 600 python bench.py --steps 20 --warmup 3 --flags 1 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_f1_$TAG.json 2>&1
for F in 0 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k 'regex:conv_ffma|conv_tc|stem_kernel|maxpool_kernel|classifier_kernel|squeeze_tokens|token_encoder|decoder_tables|pixel_decoder' \
      -s 126 -c 84 --csv --log-file gpurun_out/launches_f${F}_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --flags $F --no-cpu-baseline > gpurun_out/ncu_launch_f${F}_$TAG.log 2>&1
done
# full capture: conv_layer2_0.0 is the 20th conv launch of a forward (19 conv launches precede it)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 19 -c 1 -o gpurun_out/prof_conv_ffma_$TAG -f \
    python bench.py --steps 1 --warmup 3 --flags 0 --no-cpu-baseline > gpurun_out/ncu_full_ffma_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 15 -c 1 -o gpurun_out/prof_conv_tc_$TAG -f \
    python bench.py --steps 1 --warmup 3 --flags 1 --no-cpu-baseline > gpurun_out/ncu_full_tc_$TAG.log 2>&1
tail -3 gpurun_out/blocks_$TAG.log gpurun_out/forward_$TAG.log gpurun_out/tc_blocks_$TAG.log gpurun_out/tc_forward_$TAG.log
cat gpurun_out/bench_f0_$TAG.json | cut -c1-400
cat gpurun_out/bench_f1_$TAG.json | cut -c1-400
ls -la gpurun_out | tail -20
