timeout 900 python -m pytest tests/test_detector_set.py -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1; tail -5 gpurun_out/r2d_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --workload landmarks15 --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wvm_group_kernel -s 14 -c 7 -o gpurun_out/grp_r2d python bench.py --workload landmarks15 --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu2.log 2>&1
ls -la gpurun_out/grp_r2d.ncu-rep
