mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_tf32x3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02ah_ncu_launch.log 2>&1
grep -c '^"' gpurun_out/r02_launches_tf32x3.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:classifier_tma_kernel -s 2 -c 1 -o gpurun_out/prof_classifier_nc5_r02 -f \
    python bench.py --workload xbd1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02ah_ncu_cls.log 2>&1
tail -n 2 gpurun_out/r02ah_ncu_cls.log
