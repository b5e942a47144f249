mkdir -p gpurun_out
cp gpurun_out/launches_tf32x3_r02e.csv gpurun_out/launches_tf32x3_r02f.csv 2>/dev/null
tools/ncu_capture_instance.sh tf32x3 r02f conv_tc3_kernel 'conv_tc3_kernel<128, 3, 1, 1, 0, 1>' conv_tc3_N128_stream
tools/ncu_capture_instance.sh tf32x3 r02f conv_tc3_kernel 'conv_tc3_kernel<32, 3, 1, 1, 1, 1>' conv_tc3_N32_res
tools/ncu_capture_instance.sh tf32x3 r02f conv_tc3_kernel 'conv_tc3_kernel<128, 3, 1, 0, 1, 1>' conv_tc3_N128_ps_res
grep -h "gpu__time_duration\|kernel:" gpurun_out/r02f_ncu_full_*.txt | cut -c1-120
