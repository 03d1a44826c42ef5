#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detector_set.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-facefrontal --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -2 gpurun_out/r2v_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); print('bench', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['stage1_ms'], d['host_ms_last_step'])"
ncu --set full --clock-control none --import-source on -k regex:"wvm_group_tc_kernel" -s 7 -c 1 -o gpurun_out/grptc_r2v python bench.py --frames 64 --steps 1 --warmup 1 --no-cpu-baseline --no-facefrontal > gpurun_out/r2v_ncu.log 2>&1
ls -la gpurun_out/grptc_r2v.ncu-rep
