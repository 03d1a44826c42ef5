#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_r3n.csv python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r3n_ncu1.log 2>&1
ls -la gpurun_out/launches_r3n.csv
