#!/bin/bash
# The forward alone, timed from the plain-C host with the inputs resident in HBM (cross-check of bench.py's `value` without
# PyTorch in the process), LEVIR 64 pairs and xBD 8 x 1024^2; then smoke() on the final tree.
mkdir -p gpurun_out
X=examples/bin
OUT=gpurun_out/r02_c_host_resident.log
(timeout 15 $X/dahitra_infer --weights $X/san_w.bin --synthetic 64x256x256 --repeat 50 --resident) > $OUT 2>&1
(timeout 15 $X/dahitra_infer --weights $X/san_w_xbd.bin --variant 1 --nc 5 --synthetic 8x1024x1024 --repeat 20 --resident) >> $OUT 2>&1
cat $OUT
(timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/r02_smoke_final.log 2>&1
cat gpurun_out/r02_smoke_final.log
