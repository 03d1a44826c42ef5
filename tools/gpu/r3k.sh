#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"resize_kernel|pyrdown_kernel" -s 60 -c 5 -o gpurun_out/pyr_r3k python bench.py --workload facefrontal --frames 256 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3k_ncu.log 2>&1
ls -la gpurun_out/pyr_r3k.ncu-rep
