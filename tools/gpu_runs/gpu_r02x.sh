mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -n 8) > gpurun_out/r02x_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/r02x_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_train_bwd -s 3 -c 1 -o gpurun_out/prof_decoder_train_bwd_r02 -f \
   python tools/bench_train_decoder.py --only "level3 pair" --iters 1 > gpurun_out/r02x_ncu_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_train_fwd -s 3 -c 1 -o gpurun_out/prof_decoder_train_fwd_r02 -f \
   python tools/bench_train_decoder.py --only "level3 pair" --iters 1 > gpurun_out/r02x_ncu_fwd.log 2>&1
tail -n 2 gpurun_out/r02x_ncu_bwd.log gpurun_out/r02x_ncu_fwd.log
(timeout 600 python bench.py --no-cpu-baseline --no-parity 2>gpurun_out/r02x.err | tail -n 1) > gpurun_out/r02x_bench_default.json
cut -c1-200 gpurun_out/r02x_bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_default.json')); print(d['strong_scaling'])"
tail -n 3 gpurun_out/r02x.err
