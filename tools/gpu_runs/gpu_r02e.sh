mkdir -p gpurun_out
rm -f gpurun_out/launches_tf32x3_r02e.csv
tools/ncu_capture_instance.sh tf32x3 r02e conv_tc3_kernel 'conv_tc3_kernel<128, 3, 1, 1, false' conv_tc3_N128_stream
tools/ncu_capture_instance.sh tf32x3 r02e conv_tc3_kernel 'conv_tc3_kernel<64, 3, 1, 1, true' conv_tc3_N64_res
tools/ncu_capture_instance.sh tf32x3 r02e conv_tc3_kernel 'conv_tc3_kernel<32, 3, 1, 1, true' conv_tc3_N32_res
tools/ncu_capture_instance.sh tf32x3 r02e conv_tc3_kernel 'conv_tc3_kernel<128, 3, 1, 0, true' conv_tc3_N128_ps_res
rm -f gpurun_out/launches_tf32x3_r02e.csv.tmp
ls -la gpurun_out/ | grep r02e; du -sh gpurun_out
