#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"wvm_group_tc_kernel" -s 24 -c 8 -o gpurun_out/grptc_r2q python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/r2q_ncu.log
ls -la gpurun_out/*.ncu-rep
