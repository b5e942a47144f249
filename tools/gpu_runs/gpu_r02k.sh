mkdir -p gpurun_out
rm -f gpurun_out/launches_tf32x3_r02.csv
tools/ncu_capture_instance.sh tf32x3 r02 pixel_decoder_tc_kernel 'pixel_decoder_tc_kernel<8' pixel_decoder_tc_level3
tools/ncu_capture_instance.sh tf32x3 r02 stem_f16_kernel 'stem_f16_kernel' stem_f16
tools/ncu_capture_instance.sh tf32x3 r02 classifier_tma_kernel 'classifier_tma_kernel' classifier_tma
tools/ncu_capture_instance.sh tf32x3 r02 maxpool_split_kernel 'maxpool_split_kernel' maxpool_split
tools/ncu_capture_instance.sh tf32x3 r02 conv_tc3_kernel 'conv_tc3_kernel<32, 1, 1, 2, 1, 1>' conv_tc3_squeeze_tok
tools/ncu_capture_instance.sh tf32x3 r02 conv_tc3_kernel 'conv_tc3_kernel<128, 3, 1, 1, 0, 1>' conv_tc3_conv_layer2_0.0
(timeout 900 python -m pytest tests/test_gpu_forward.py -q -x -k "c_abi_only" 2>&1 | tail -3) > gpurun_out/r02k_cabi.log 2>&1
grep -h "gpu__time_duration\|kernel:" gpurun_out/r02_ncu_full_*.txt | cut -c1-110; cat gpurun_out/r02k_cabi.log
