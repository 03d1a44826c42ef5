#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"wvm_group_tc_kernel" -s 1 -c 1 -o gpurun_out/grptc_r2x python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2x_ncu.log 2>&1
ls -la gpurun_out/grptc_r2x.ncu-rep
